// Definition-file readers (pure host code, no CUDA).
//
// Formats are the reference's: the Columbus7 fixed-column `intcfl` and the free "default"
// format (reference source/IC/IntCoordSet.cpp:11-120, documented in intcoord.md:23-33), and
// the SAP term list (source/polynomial/SAPSet.cpp:89-109, source/polynomial/SAP.cpp:42-76).
// Behaviour is matched, including the quirks a drop-in has to keep:
//   * a last line without a trailing newline is ignored (getline / good() idiom, :19-20,:64-65);
//   * Columbus7: the loop stops at the first line that is not STRE/BEND/TORS/OUT (:51);
//   * default format: any format string other than "Columbus7" selects it (:56);
//   * coefficients are normalised per internal coordinate after reading (:118-119).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "common.h"

namespace tchem_b200 {

namespace {

const char * const kTypeNames[TCHEM_NTYPES] = {
    "dummy", "stretching", "bending", "cosbending", "torsion", "sintorsion", "costorsion",
    "torsion2", "sintorsion2", "costorsion2", "pxtorsion2", "pytorsion2", "OutOfPlane", "sinoop"};
const int kTypeAtoms[TCHEM_NTYPES] = {0, 2, 3, 3, 4, 4, 4, 5, 5, 5, 5, 5, 4, 4};

// lines as std::getline + `if (!ifs.good()) break` sees them
std::vector<std::string> complete_lines(const std::string & text) {
    std::vector<std::string> lines;
    size_t pos = 0;
    while (true) {
        size_t nl = text.find('\n', pos);
        if (nl == std::string::npos) break;      // unterminated tail is dropped
        lines.push_back(text.substr(pos, nl - pos));
        pos = nl + 1;
    }
    return lines;
}

std::vector<std::string> tokens(const std::string & line) {
    std::vector<std::string> out;
    std::istringstream iss(line);
    std::string tok;
    while (iss >> tok) out.push_back(tok);
    return out;
}

std::string field(const std::string & line, size_t pos, size_t len) {
    if (pos >= line.size()) return std::string();
    return line.substr(pos, len);
}

// the reference accepts a trailing token as `min` when it matches -?\d+\.?\d*
bool looks_like_min(const std::string & s) {
    size_t i = 0;
    if (i < s.size() && s[i] == '-') i++;
    size_t d0 = i;
    while (i < s.size() && std::isdigit((unsigned char)s[i])) i++;
    if (i == d0) return false;
    if (i < s.size() && s[i] == '.') i++;
    while (i < s.size() && std::isdigit((unsigned char)s[i])) i++;
    return i == s.size();
}

int atom_from(const std::string & s, bool & ok) {
    try {
        size_t used = 0;
        unsigned long v = std::stoul(s, &used);
        if (v == 0) { ok = false; return 0; }
        return (int)(v - 1);
    } catch (...) { ok = false; return 0; }
}

void normalise(std::vector<tchem_invdisp_t> & terms, int n_ic) {
    // IntCoord::normalize, source/IC/IntCoord.cpp:26-31: c <- c / sqrt(sum c^2)
    std::vector<double> norm2(n_ic, 0.0);
    for (auto & t : terms) norm2[t.ic] += t.coeff * t.coeff;
    for (auto & n : norm2) n /= std::sqrt(n);
    for (auto & t : terms) t.coeff /= norm2[t.ic];
}

int parse_columbus7(const std::vector<std::string> & lines, std::vector<tchem_invdisp_t> & terms, int & n_ic) {
    for (size_t li = 1; li < lines.size(); li++) {          // line 0 is "TEXAS"
        const std::string & line = lines[li];
        bool new_ic = !line.empty() && line[0] == 'K';
        tchem_invdisp_t t;
        std::memset(&t, 0, sizeof(t));
        t.coeff = 1.0;
        t.min = -M_PI;
        std::string cf = field(line, 10, 10);
        bool ok = true;
        if (cf != std::string(10, ' ')) {
            try { t.coeff = std::stod(cf); } catch (...) { return set_error(TCHEM_EINVAL, "tchem_ic_parse: bad Columbus7 coefficient field: " + line); }
        }
        std::string key = field(line, 20, 4);
        // atom fields are fixed-width floats such as "7."; BEND lists the vertex last and OUT the
        // centre last, both are permuted into tchem order
        if (key == "STRE") {
            t.type = TCHEM_STRETCHING; t.natoms = 2;
            t.atoms[0] = atom_from(field(line, 28, 5), ok); t.atoms[1] = atom_from(field(line, 34, 9), ok);
        } else if (key == "BEND") {
            t.type = TCHEM_BENDING; t.natoms = 3;
            t.atoms[0] = atom_from(field(line, 28, 6), ok); t.atoms[1] = atom_from(field(line, 45, 9), ok);
            t.atoms[2] = atom_from(field(line, 35, 9), ok);
        } else if (key == "TORS") {
            t.type = TCHEM_TORSION; t.natoms = 4;
            t.atoms[0] = atom_from(field(line, 28, 6), ok); t.atoms[1] = atom_from(field(line, 35, 9), ok);
            t.atoms[2] = atom_from(field(line, 45, 9), ok); t.atoms[3] = atom_from(field(line, 55, 9), ok);
        } else if (key.substr(0, 3) == "OUT") {
            t.type = TCHEM_OUTOFPLANE; t.natoms = 4;
            t.atoms[0] = atom_from(field(line, 28, 6), ok); t.atoms[1] = atom_from(field(line, 55, 9), ok);
            t.atoms[2] = atom_from(field(line, 35, 9), ok); t.atoms[3] = atom_from(field(line, 45, 9), ok);
        } else break;
        if (!ok) return set_error(TCHEM_EINVAL, "tchem_ic_parse: bad Columbus7 atom field: " + line);
        if (new_ic) n_ic++;
        if (n_ic == 0) return set_error(TCHEM_EINVAL, "tchem_ic_parse: Columbus7 term before the first 'K' line");
        t.ic = n_ic - 1;
        terms.push_back(t);
    }
    return TCHEM_OK;
}

int parse_default(const std::vector<std::string> & lines, std::vector<tchem_invdisp_t> & terms, int & n_ic) {
    for (const std::string & line : lines) {
        std::vector<std::string> tok = tokens(line);
        size_t k = 0;
        bool new_ic = line.substr(0, 6) != "      ";
        if (new_ic) k++;                                     // leading index token is discarded
        if (tok.size() < k + 2)
            return set_error(TCHEM_EINVAL, "Error during reading internal coordinate definition: incomplete line: " + line);
        tchem_invdisp_t t;
        std::memset(&t, 0, sizeof(t));
        try { t.coeff = std::stod(tok[k++]); }
        catch (...) { return set_error(TCHEM_EINVAL, "Error during reading internal coordinate definition: bad coefficient: " + line); }
        const std::string & name = tok[k++];
        t.type = -1;
        for (int i = 0; i < TCHEM_NTYPES; i++) if (name == kTypeNames[i]) t.type = i;
        if (t.type < 0)
            return set_error(TCHEM_EINVAL, "Error during reading internal coordinate definition: unsupported internal coordinate type: " + name);
        t.natoms = kTypeAtoms[t.type];
        if (tok.size() < k + (size_t)t.natoms)
            return set_error(TCHEM_EINVAL, "Error during reading internal coordinate definition: too few atoms: " + line);
        bool ok = true;
        for (int a = 0; a < t.natoms; a++) t.atoms[a] = atom_from(tok[k++], ok);
        if (!ok) return set_error(TCHEM_EINVAL, "Error during reading internal coordinate definition: bad atom index: " + line);
        t.min = -M_PI;
        if (k < tok.size() && looks_like_min(tok[k])) t.min = std::stod(tok[k]);
        if (new_ic) n_ic++;
        if (n_ic == 0) return set_error(TCHEM_EINVAL, "Error during reading internal coordinate definition: continuation line before the first coordinate");
        t.ic = n_ic - 1;
        terms.push_back(t);
    }
    return TCHEM_OK;
}

int read_file(const char * file, std::string & text) {
    std::ifstream ifs(file);
    if (!ifs.good()) return set_error(TCHEM_EFILE, std::string(file ? file : "(null)"));
    std::stringstream ss;
    ss << ifs.rdbuf();
    text = ss.str();
    return TCHEM_OK;
}

int export_terms(const std::vector<tchem_invdisp_t> & v, int nic, tchem_invdisp_t * terms, int capacity,
                 int * n_terms, int * n_ic) {
    if (n_terms) *n_terms = (int)v.size();
    if (n_ic) *n_ic = nic;
    if (terms) std::copy(v.begin(), v.begin() + std::min<size_t>(v.size(), capacity < 0 ? 0 : capacity), terms);
    return TCHEM_OK;
}

} // namespace

int parse_ic_text(const std::string & format, const std::string & text, std::vector<tchem_invdisp_t> & terms, int & n_ic) {
    terms.clear();
    n_ic = 0;
    std::vector<std::string> lines = complete_lines(text);
    int rc = format == "Columbus7" ? parse_columbus7(lines, terms, n_ic) : parse_default(lines, terms, n_ic);
    if (rc != TCHEM_OK) return rc;
    normalise(terms, n_ic);
    return TCHEM_OK;
}

// One monomial per line: whitespace separated "irreducible,index" tokens (1-based), a bare '#'
// token ends the line, no tokens = the constant term. coords_ are sorted descending like
// SAP::SAP (source/polynomial/SAP.cpp:51).
int parse_sap_text(const std::string & text, std::vector<int32_t> & term_ptr, std::vector<int32_t> & coords) {
    term_ptr.assign(1, 0);
    coords.clear();
    for (const std::string & line : complete_lines(text)) {
        std::vector<std::pair<int32_t, int32_t>> cs;
        for (const std::string & tok : tokens(line)) {
            if (tok == "#") break;
            size_t comma = tok.find(',');
            try {
                if (comma == std::string::npos) throw 0;
                int32_t a = (int32_t)std::stoul(tok.substr(0, comma)), b = (int32_t)std::stoul(tok.substr(comma + 1));
                if (a < 1 || b < 1) throw 0;
                cs.emplace_back(a - 1, b - 1);
            } catch (...) { return set_error(TCHEM_EINVAL, "tchem_sap_parse: bad token '" + tok + "'"); }
        }
        std::sort(cs.begin(), cs.end(), std::greater<std::pair<int32_t, int32_t>>());
        for (auto & c : cs) { coords.push_back(c.first); coords.push_back(c.second); }
        term_ptr.push_back((int32_t)(coords.size() / 2));
    }
    return TCHEM_OK;
}

} // namespace tchem_b200

using namespace tchem_b200;

extern "C" {

int tchem_ic_parse_text(const char * format, const char * text, tchem_invdisp_t * terms, int capacity,
                        int * n_terms, int * n_ic) {
    if (!format || !text) return set_error(TCHEM_EINVAL, "tchem_ic_parse_text: null argument");
    std::vector<tchem_invdisp_t> v;
    int nic = 0;
    int rc = parse_ic_text(format, text, v, nic);
    if (rc != TCHEM_OK) return rc;
    return export_terms(v, nic, terms, capacity, n_terms, n_ic);
}

int tchem_ic_parse_file(const char * format, const char * file, tchem_invdisp_t * terms, int capacity,
                        int * n_terms, int * n_ic) {
    if (!format || !file) return set_error(TCHEM_EINVAL, "tchem_ic_parse_file: null argument");
    std::string text;
    int rc = read_file(file, text);
    if (rc != TCHEM_OK) return rc;
    return tchem_ic_parse_text(format, text.c_str(), terms, capacity, n_terms, n_ic);
}

int tchem_sap_parse_text(const char * text, int32_t * term_ptr, int capacity_terms, int32_t * coords,
                         int capacity_coords, int * n_terms, int * n_coords) {
    if (!text) return set_error(TCHEM_EINVAL, "tchem_sap_parse_text: null argument");
    std::vector<int32_t> tp, cs;
    int rc = parse_sap_text(text, tp, cs);
    if (rc != TCHEM_OK) return rc;
    int nt = (int)tp.size() - 1, nc = (int)cs.size() / 2;
    if (n_terms) *n_terms = nt;
    if (n_coords) *n_coords = nc;
    if (term_ptr && capacity_terms >= nt) std::copy(tp.begin(), tp.end(), term_ptr);
    if (coords && capacity_coords >= nc) std::copy(cs.begin(), cs.end(), coords);
    return TCHEM_OK;
}

int tchem_sap_parse_file(const char * file, int32_t * term_ptr, int capacity_terms, int32_t * coords,
                         int capacity_coords, int * n_terms, int * n_coords) {
    if (!file) return set_error(TCHEM_EINVAL, "tchem_sap_parse_file: null argument");
    std::string text;
    int rc = read_file(file, text);
    if (rc != TCHEM_OK) return rc;
    return tchem_sap_parse_text(text.c_str(), term_ptr, capacity_terms, coords, capacity_coords, n_terms, n_coords);
}

} // extern "C"
