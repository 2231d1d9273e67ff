// C ABI glue: error state, table life cycle, q/J/K entry points (include/tchem_b200.h).
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <utility>
#include <string>
#include <vector>

#include "ic_table.h"

namespace tchem_b200 {

namespace {
thread_local std::string g_error;
std::atomic<int64_t> g_launches{0};
std::mutex g_build_mutex;
}

int set_error(int code, const std::string & msg) { g_error = msg; return code; }

int ensure_max_dynamic_smem(const void * fn, int device, int max_smem_optin) {
    static std::mutex m;
    static std::set<std::pair<const void *, int>> done;
    std::lock_guard<std::mutex> lock(m);
    if (done.count({fn, device})) return TCHEM_OK;
    // (the opt-in maximum covers static + dynamic shared memory)
    cudaFuncAttributes fa;
    TC_CUDA(cudaFuncGetAttributes(&fa, fn));
    TC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin - (int)fa.sharedSizeBytes));
    done.insert({fn, device});
    return TCHEM_OK;
}
void keep_async_pool(int device) {
    static std::mutex m;
    static std::set<int> done;
    std::lock_guard<std::mutex> lock(m);
    if (done.count(device)) return;
    done.insert(device);
    const char * e = getenv("TCHEM_KEEP_WORKSPACE");
    if (e && atoi(e) == 0) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) != cudaSuccess) { cudaGetLastError(); return; }
    uint64_t threshold = UINT64_MAX;
    if (cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold) != cudaSuccess) cudaGetLastError();
}
namespace {
std::mutex g_log_mutex;
std::map<std::string, int64_t> & kernel_log() { static std::map<std::string, int64_t> m; return m; }
}
void count_launch(const char * kernel, int n) {
    g_launches += n;
    if (kernel) { std::lock_guard<std::mutex> lock(g_log_mutex); kernel_log()[kernel] += n; }
}
std::string kernel_log_text() {
    std::lock_guard<std::mutex> lock(g_log_mutex);
    std::string out;
    for (const auto & kv : kernel_log()) out += kv.first + "=" + std::to_string(kv.second) + ";";
    return out;
}

void release_grad_program(const tchem_ic_table * t);
void release_dense_program(const tchem_ic_table * t);
void release_sparse_program(const tchem_ic_table * t);
void release_jit(const tchem_ic_table * t);
extern thread_local int64_t g_batch_hint;
int launch_ic_eval(const tchem_ic_table * t, int mode, const double * r, int64_t B,
                   double * q, double * J, double * K, cudaStream_t stream);

// build + upload the program of one mode on first use
int ensure_program(tchem_ic_table * t, int mode) {
    std::lock_guard<std::mutex> lock(g_build_mutex);
    if (t->dev_blob[mode]) return TCHEM_OK;
    HostProgram hp;
    int rc = TCHEM_OK;
    bool grouped = false;
    if (mode == MODE_QJ && t->cartdim > 36) {
        // mid-size molecules (beyond the specialised kernel's range): q + J runs on the warp-group
        // kernel - G warps share a tile's staged coordinates, 12 warps per SM. The compact capacity
        // per warp is what is left of the shared memory (TCHEM_GROUP_G=0 disables, 2 / 4 / 6 force G)
        static const int g_env = [] { const char * e = getenv("TCHEM_GROUP_G"); return e ? atoi(e) : 6; }();
        const int G = (g_env == 2 || g_env == 4 || g_env == 6) ? g_env : 0;
        if (G) {
            const int ngroups = 12 / G;
            // per group: cartdim coordinate rows; per warp `cap` compact rows (kPad doubles each) and two
            // line buffers of max_rows * cartdim doubles. Rows per chunk: 2, 3 or 4, whichever gives the
            // cheapest schedule - the group's warps take chunks round-robin, a round lasts as long as
            // its most expensive chunk (arithmetic ~ compact entries, write-out ~ 0.15 per element:
            // ratio read off the 20-atom profile). TCHEM_GROUP_ROWS forces a value.
            static const int rows_env = [] { const char * e = getenv("TCHEM_GROUP_ROWS"); return e ? atoi(e) : 0; }();
            const int group_doubles = ((t->max_smem_sm - 1024 - 12 * 1024) / ngroups) / (int)sizeof(double);
            double best_cost = 0.0;
            for (int max_rows = 2; max_rows <= 4; max_rows++) {
                if (rows_env > 0 && max_rows != std::min(std::max(rows_env, 2), 4)) continue;
                const int line = 2 * ((max_rows * t->cartdim + 1) & ~1);
                const int cap = (group_doubles - t->cartdim * kPad - G * line) / (G * kPad);
                if (cap < 8) continue;
                HostProgram cand;
                if (build_program(t->terms, t->intdim, t->cartdim, mode, cand, cap, max_rows) != TCHEM_OK || cand.cap > cap) continue;
                bool split = false;
                for (const ChunkDesc & ch : cand.chunks) split = split || ch.accumulate;
                if (split) continue;
                double cost = 0.0, round_max = 0.0;
                for (size_t c = 0; c < cand.chunks.size(); c++) {
                    const ChunkDesc & ch = cand.chunks[c];
                    int entries = ch.nrows;
                    for (int k = ch.prim_begin; k < ch.prim_end; k++) entries += prim_entries(cand.prims[k].natoms, mode);
                    round_max = std::max(round_max, entries + 0.15 * ch.nrows * t->cartdim);
                    if (c % G == (size_t)G - 1 || c + 1 == cand.chunks.size()) { cost += round_max; round_max = 0.0; }
                }
                if (!grouped || cost < best_cost) { hp = cand; best_cost = cost; grouped = true; }
            }
            if (grouped) t->group_warps = G;
        }
    }
    if (!grouped) rc = build_program(t->terms, t->intdim, t->cartdim, mode, hp);
    if (rc != TCHEM_OK) return rc;
    auto a16 = [](size_t x) { return (x + 15) & ~size_t(15); };
    const size_t o_prims = 0;
    const size_t o_chunks = o_prims + a16(hp.prims.size() * sizeof(PrimDesc));
    const size_t o_map = o_chunks + a16(hp.chunks.size() * sizeof(ChunkDesc));
    const size_t o_qsrcs = o_map + a16(hp.map.size() * sizeof(uint32_t));
    const size_t o_emap = o_qsrcs + a16(hp.qsrcs.size() * sizeof(Src));
    const size_t o_srcs = o_emap + a16(hp.emap.size() * sizeof(ElemMap));
    const size_t o_coefs = o_srcs + a16(hp.psrcs.size() * sizeof(uint32_t));
    const size_t o_runs = o_coefs + a16(hp.coefs.size() * sizeof(double));
    const size_t total = o_runs + a16(hp.runs.size() * sizeof(ZeroRun)) + 16;
    std::vector<unsigned char> blob(total, 0);
    std::memcpy(blob.data() + o_runs, hp.runs.data(), hp.runs.size() * sizeof(ZeroRun));
    std::memcpy(blob.data() + o_prims, hp.prims.data(), hp.prims.size() * sizeof(PrimDesc));
    std::memcpy(blob.data() + o_chunks, hp.chunks.data(), hp.chunks.size() * sizeof(ChunkDesc));
    std::memcpy(blob.data() + o_map, hp.map.data(), hp.map.size() * sizeof(uint32_t));
    std::memcpy(blob.data() + o_qsrcs, hp.qsrcs.data(), hp.qsrcs.size() * sizeof(Src));
    std::memcpy(blob.data() + o_emap, hp.emap.data(), hp.emap.size() * sizeof(ElemMap));
    std::memcpy(blob.data() + o_srcs, hp.psrcs.data(), hp.psrcs.size() * sizeof(uint32_t));
    std::memcpy(blob.data() + o_coefs, hp.coefs.data(), hp.coefs.size() * sizeof(double));
    DeviceGuard guard(t->device);
    if (!guard.ok) return set_error(TCHEM_ECUDA, "cudaSetDevice failed");
    void * dev = nullptr;
    TC_CUDA(cudaMalloc(&dev, total));
    cudaError_t e = cudaMemcpy(dev, blob.data(), total, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(dev); return set_error(TCHEM_ECUDA, cudaGetErrorString(e)); }
    Program & p = t->prog[mode];
    unsigned char * d = static_cast<unsigned char *>(dev);
    p.prims = reinterpret_cast<const PrimDesc *>(d + o_prims);
    p.chunks = reinterpret_cast<const ChunkDesc *>(d + o_chunks);
    p.map = reinterpret_cast<const uint32_t *>(d + o_map);
    p.qsrcs = reinterpret_cast<const Src *>(d + o_qsrcs);
    p.nqsrcs = (int)hp.qsrcs.size();
    p.nmap = (int)hp.map.size();
    p.split_rows = 0;
    for (const ChunkDesc & ch : hp.chunks) if (ch.accumulate) p.split_rows = 1;
    p.emap = reinterpret_cast<const ElemMap *>(d + o_emap);
    p.psrcs = reinterpret_cast<const uint4 *>(d + o_srcs);
    p.coefs = reinterpret_cast<const double *>(d + o_coefs);
    p.ncoefs = (int)hp.coefs.size();
    p.runs = reinterpret_cast<const ZeroRun *>(d + o_runs);
    p.nprims = (int)hp.prims.size();
    p.nchunks = (int)hp.chunks.size();
    p.cap = hp.cap;
    p.intdim = t->intdim;
    p.cartdim = t->cartdim;
    p.line_elems = 0;
    if (mode == MODE_QJ && grouped) {
        // multi-row chunks of the warp-group kernel: J pieces up to 1024 doubles go out through the
        // per-warp line buffer (TCHEM_LINE=0 disables). One-row chunks gain nothing from it (measured).
        static const int on = [] { const char * e = getenv("TCHEM_LINE"); return e ? atoi(e) : 1; }();
        int widest = 0;
        for (const ChunkDesc & ch : hp.chunks) widest = std::max(widest, ch.nrows * t->cartdim);
        if (on && widest <= 1024) p.line_elems = 2 * ((widest + 1) & ~1);      // two buffers (double-buffered write-out)
    }
    // small J-mode tables are staged per block in shared memory: under a saturated store stream the
    // L2 latency of table reads is what the write-out waits for (TCHEM_STAGE_SPANS=0 disables)
    {
        static const int limit = [] { const char * e = getenv("TCHEM_STAGE_SPANS"); return e ? atoi(e) : 48 * 1024; }();
        const size_t bytes = (o_runs + a16(hp.runs.size() * sizeof(ZeroRun))) - o_emap;
        p.span_bytes = (mode != MODE_QJK && !grouped && bytes <= (size_t)limit) ? (int32_t)bytes : 0;
    }
    t->dev_blob[mode] = dev;
    return TCHEM_OK;
}

} // namespace tchem_b200

using namespace tchem_b200;

// ---- host pipeline for the *_host entry points ------------------------------------------
struct HostPipe {
    static constexpr int kSlots = 3;
    cudaStream_t stream[kSlots] = {};
    double * d_r[kSlots] = {};
    double * d_q[kSlots] = {};
    double * d_J[kSlots] = {};
    double * d_K[kSlots] = {};
    size_t cap_r = 0, cap_q = 0, cap_J = 0, cap_K = 0;   // bytes per slot
    bool streams_ok = false;
    void release() {
        for (int s = 0; s < kSlots; s++) {
            cudaFree(d_r[s]); cudaFree(d_q[s]); cudaFree(d_J[s]); cudaFree(d_K[s]);
            d_r[s] = d_q[s] = d_J[s] = d_K[s] = nullptr;
            if (streams_ok) cudaStreamDestroy(stream[s]);
        }
        streams_ok = false;
        cap_r = cap_q = cap_J = cap_K = 0;
    }
};

namespace {
int grow(double ** slots, size_t & cap, size_t need) {
    if (need <= cap) return TCHEM_OK;
    for (int s = 0; s < HostPipe::kSlots; s++) {
        cudaFree(slots[s]);
        slots[s] = nullptr;
        TC_CUDA(cudaMalloc(reinterpret_cast<void **>(&slots[s]), need));
    }
    cap = need;
    return TCHEM_OK;
}
}

extern "C" {

const char * tchem_last_error(void) { return g_error.c_str(); }

int tchem_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int64_t tchem_launch_count(void) { return g_launches.load(); }

int64_t tchem_kernel_log(char * buf, int64_t capacity) {
    const std::string text = kernel_log_text();
    if (buf && capacity > 0) {
        const size_t n = std::min<size_t>(text.size(), (size_t)capacity - 1);
        std::memcpy(buf, text.data(), n);
        buf[n] = 0;
    }
    return (int64_t)text.size() + 1;
}

int tchem_ic_table_create(const tchem_invdisp_t * terms, int n_terms, int n_ic, int n_atoms,
                          int device, tchem_ic_table ** out) {
    if (!terms || !out || n_terms <= 0 || n_ic <= 0) return set_error(TCHEM_EINVAL, "tchem_ic_table_create: empty definition");
    int max_atom = -1;
    for (int i = 0; i < n_terms; i++)
        for (int a = 0; a < terms[i].natoms && a < 5; a++) max_atom = std::max(max_atom, terms[i].atoms[a]);
    if (n_atoms <= 0) n_atoms = max_atom + 1;
    if (n_atoms <= max_atom) return set_error(TCHEM_EINVAL, "tchem_ic_table_create: atom index exceeds n_atoms");
    if (n_atoms <= 0) n_atoms = 1;
    tchem_ic_table * t = new tchem_ic_table();
    t->device = device;
    t->intdim = n_ic;
    t->natoms = n_atoms;
    t->cartdim = 3 * n_atoms;
    t->terms.assign(terms, terms + n_terms);
    for (const auto & d : t->terms) if (d.natoms == 5) t->has5 = true;
    std::memset(t->prog, 0, sizeof(t->prog));
    // validate once on the host (cheap mode) so that definition errors surface here
    HostProgram hp;
    int rc = build_program(t->terms, t->intdim, t->cartdim, MODE_VALUE, hp);
    if (rc != TCHEM_OK) { delete t; return rc; }
    int ndev = tchem_device_count();
    if (device < 0 || device >= ndev) {
        delete t;
        return set_error(TCHEM_ECUDA, "tchem_ic_table_create: no such CUDA device (this library has no CPU path)");
    }
    cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&t->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&t->max_smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
    *out = t;
    return TCHEM_OK;
}

void tchem_ic_table_destroy(tchem_ic_table * t) {
    if (!t) return;
    {
        DeviceGuard guard(t->device);
        for (int m = 0; m < MODE_COUNT; m++) cudaFree(t->dev_blob[m]);
        release_grad_program(t);
        release_dense_program(t);
        release_sparse_program(t);
        release_jit(t);
        if (t->pipe) { t->pipe->release(); delete t->pipe; }
    }
    delete t;
}
int tchem_ic_table_intdim(const tchem_ic_table * t) { return t ? t->intdim : 0; }
int tchem_ic_table_cartdim(const tchem_ic_table * t) { return t ? t->cartdim : 0; }
int tchem_ic_table_device(const tchem_ic_table * t) { return t ? t->device : -1; }

static int pick_mode(double * J, double * K, int value_path, int & mode) {
    if (value_path) {
        if (J || K) return set_error(TCHEM_EINVAL, "tchem_ic_eval: the value path (IntCoordSet::operator()) has no J / K");
        mode = MODE_VALUE;
    } else mode = K ? MODE_QJK : MODE_QJ;
    return TCHEM_OK;
}

int tchem_ic_eval(const tchem_ic_table * tc, const double * r, int64_t B, double * q, double * J, double * K,
                  int value_path, tchem_stream_t stream) {
    tchem_ic_table * t = const_cast<tchem_ic_table *>(tc);
    if (!t) return set_error(TCHEM_EINVAL, "tchem_ic_eval: null table");
    if (B < 0) return set_error(TCHEM_EINVAL, "tchem_ic_eval: negative batch");
    if (B == 0) return TCHEM_OK;
    if (!r) return set_error(TCHEM_EINVAL, "tchem_ic_eval: r is null");
    if (!q && !J && !K) return set_error(TCHEM_EINVAL, "tchem_ic_eval: no output requested");
    int mode;
    int rc = pick_mode(J, K, value_path, mode);
    if (rc != TCHEM_OK) return rc;
    rc = ensure_program(t, mode);
    if (rc != TCHEM_OK) return rc;
    DeviceGuard guard(t->device);
    if (!guard.ok) return set_error(TCHEM_ECUDA, "cudaSetDevice failed");
    return launch_ic_eval(t, mode, r, B, q, J, K, static_cast<cudaStream_t>(stream));
}

int tchem_ic_program_export(const tchem_invdisp_t * terms, int n_terms, int n_ic, int cartdim, int mode,
                            int64_t * counts, void * prims, void * chunks, uint32_t * map, void * qsrcs,
                            void * emap, void * srcs, void * runs, double * coefs) {
    if (!terms || !counts || n_terms <= 0 || n_ic <= 0 || mode < 0 || mode >= MODE_COUNT)
        return set_error(TCHEM_EINVAL, "tchem_ic_program_export: bad arguments");
    std::vector<tchem_invdisp_t> v(terms, terms + n_terms);
    HostProgram hp;
    int rc = build_program(v, n_ic, cartdim, mode, hp);
    if (rc != TCHEM_OK) return rc;
    counts[0] = (int64_t)hp.prims.size(); counts[1] = (int64_t)hp.chunks.size(); counts[2] = (int64_t)hp.map.size();
    counts[3] = (int64_t)hp.qsrcs.size(); counts[4] = (int64_t)hp.emap.size(); counts[5] = (int64_t)hp.psrcs.size();
    counts[8] = (int64_t)hp.coefs.size();
    counts[6] = hp.cap;
    counts[7] = (int64_t)hp.runs.size();
    if (runs) std::memcpy(runs, hp.runs.data(), hp.runs.size() * sizeof(ZeroRun));
    if (prims) std::memcpy(prims, hp.prims.data(), hp.prims.size() * sizeof(PrimDesc));
    if (chunks) std::memcpy(chunks, hp.chunks.data(), hp.chunks.size() * sizeof(ChunkDesc));
    if (map) std::memcpy(map, hp.map.data(), hp.map.size() * sizeof(uint32_t));
    if (qsrcs) std::memcpy(qsrcs, hp.qsrcs.data(), hp.qsrcs.size() * sizeof(Src));
    if (emap) std::memcpy(emap, hp.emap.data(), hp.emap.size() * sizeof(ElemMap));
    if (srcs) std::memcpy(srcs, hp.psrcs.data(), hp.psrcs.size() * sizeof(uint32_t));
    if (coefs) std::memcpy(coefs, hp.coefs.data(), hp.coefs.size() * sizeof(double));
    return TCHEM_OK;
}

int tchem_ic_eval_timed(const tchem_ic_table * tc, const double * r, int64_t B, double * q, double * J,
                        double * K, int value_path, int iters, tchem_stream_t stream, float * ms_per_launch) {
    if (iters < 1 || !ms_per_launch) return set_error(TCHEM_EINVAL, "tchem_ic_eval_timed: bad arguments");
    DeviceGuard guard(tchem_ic_table_device(tc));
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaEvent_t e0, e1;
    TC_CUDA(cudaEventCreate(&e0));
    TC_CUDA(cudaEventCreate(&e1));
    TC_CUDA(cudaEventRecord(e0, s));
    int rc = TCHEM_OK;
    for (int i = 0; i < iters && rc == TCHEM_OK; i++) rc = tchem_ic_eval(tc, r, B, q, J, K, value_path, stream);
    cudaEventRecord(e1, s);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc != TCHEM_OK) return rc;
    if (e != cudaSuccess) return set_error(TCHEM_ECUDA, cudaGetErrorString(e));
    *ms_per_launch = ms / iters;
    return TCHEM_OK;
}

// Host buffers in, host buffers out. The batch is cut into pieces; piece i uses slot i % 3
// (own stream + device buffers), so the upload of piece i+1 and the download of piece i-1
// overlap the kernel of piece i. Pinned user buffers are DMA'd directly; pageable ones go
// through the driver's staging.
int tchem_ic_eval_host(const tchem_ic_table * tc, const double * r, int64_t B, double * q, double * J,
                       double * K, int value_path) {
    tchem_ic_table * t = const_cast<tchem_ic_table *>(tc);
    if (!t) return set_error(TCHEM_EINVAL, "tchem_ic_eval_host: null table");
    if (B < 0) return set_error(TCHEM_EINVAL, "tchem_ic_eval_host: negative batch");
    if (B == 0) return TCHEM_OK;
    if (!r) return set_error(TCHEM_EINVAL, "tchem_ic_eval_host: r is null");
    if (!q && !J && !K) return set_error(TCHEM_EINVAL, "tchem_ic_eval_host: no output requested");
    int mode;
    int rc = pick_mode(J, K, value_path, mode);
    if (rc != TCHEM_OK) return rc;
    rc = ensure_program(t, mode);
    if (rc != TCHEM_OK) return rc;
    DeviceGuard guard(t->device);
    if (!guard.ok) return set_error(TCHEM_ECUDA, "cudaSetDevice failed");

    const size_t cd = t->cartdim, n = t->intdim;
    const size_t out_bytes = 8 * ((q ? n : 0) + (J ? n * cd : 0) + (K ? n * cd * cd : 0));
    // ~48 MiB of results per piece, a multiple of the 32-geometry tile
    int64_t piece = (int64_t)std::max<size_t>(32, ((size_t(48) << 20) / std::max<size_t>(out_bytes, 1)) / 32 * 32);
    piece = std::min<int64_t>(piece, (B + 31) / 32 * 32);
    std::lock_guard<std::mutex> lock(t->pipe_mutex);      // per table: calls on different tables (devices) overlap
    if (!t->pipe) t->pipe = new HostPipe();
    HostPipe & hp = *t->pipe;
    if (!hp.streams_ok) {
        for (int s = 0; s < HostPipe::kSlots; s++) TC_CUDA(cudaStreamCreateWithFlags(&hp.stream[s], cudaStreamNonBlocking));
        hp.streams_ok = true;
    }
    if ((rc = grow(hp.d_r, hp.cap_r, piece * cd * 8)) != TCHEM_OK) return rc;
    if (q && (rc = grow(hp.d_q, hp.cap_q, piece * n * 8)) != TCHEM_OK) return rc;
    if (J && (rc = grow(hp.d_J, hp.cap_J, piece * n * cd * 8)) != TCHEM_OK) return rc;
    if (K && (rc = grow(hp.d_K, hp.cap_K, piece * n * cd * cd * 8)) != TCHEM_OK) return rc;

    int slot = 0;
    struct HintScope { HintScope(int64_t b) { g_batch_hint = b; } ~HintScope() { g_batch_hint = 0; } } hint(B);
    for (int64_t b0 = 0; b0 < B; b0 += piece, slot = (slot + 1) % HostPipe::kSlots) {
        const int64_t nb = std::min<int64_t>(piece, B - b0);
        cudaStream_t s = hp.stream[slot];
        TC_CUDA(cudaMemcpyAsync(hp.d_r[slot], r + b0 * cd, nb * cd * 8, cudaMemcpyHostToDevice, s));
        rc = launch_ic_eval(t, mode, hp.d_r[slot], nb, q ? hp.d_q[slot] : nullptr, J ? hp.d_J[slot] : nullptr,
                            K ? hp.d_K[slot] : nullptr, s);
        if (rc != TCHEM_OK) {                      // no copy into the caller's buffers may stay pending
            for (int k = 0; k < HostPipe::kSlots; k++) cudaStreamSynchronize(hp.stream[k]);
            return rc;
        }
        if (q) TC_CUDA(cudaMemcpyAsync(q + b0 * n, hp.d_q[slot], nb * n * 8, cudaMemcpyDeviceToHost, s));
        if (J) TC_CUDA(cudaMemcpyAsync(J + b0 * n * cd, hp.d_J[slot], nb * n * cd * 8, cudaMemcpyDeviceToHost, s));
        if (K) TC_CUDA(cudaMemcpyAsync(K + b0 * n * cd * cd, hp.d_K[slot], nb * n * cd * cd * 8, cudaMemcpyDeviceToHost, s));
    }
    for (int s = 0; s < HostPipe::kSlots; s++) TC_CUDA(cudaStreamSynchronize(hp.stream[s]));
    return TCHEM_OK;
}

} // extern "C"
