// Per-primitive geometry of the invariant displacements, in registers.
//
// Everything is templated on a scalar type S that is either `double` or a forward-mode dual
// number. The reference has no analytic second derivative: it differentiates its analytic
// Wilson-B expression with autograd (source/IC/InvDisp.cpp:859-868) and, for the 5-atom
// torsion2 family, differentiates the value expression once or twice (:415-570, :893-1065).
// Here the same expressions are differentiated in registers: Dual carries one tangent
// direction, Dual2 two directions plus the mixed second-order term.
#pragma once
#ifndef __CUDACC_RTC__          // NVRTC (specialised kernels, jit.cu) has the math built-ins without headers
#include <cuda_runtime.h>
#include <math.h>
#endif

namespace tchem_b200 {

#define TC_PI 3.14159265358979323846

// ---------------------------------------------------------------------------------------
// first-order dual number  v + d e,  e^2 = 0
// ---------------------------------------------------------------------------------------
struct Dual {
    double v, d;
};
__device__ __forceinline__ Dual mk(double v, double d) { Dual r; r.v = v; r.d = d; return r; }
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return mk(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return mk(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ Dual operator-(Dual a) { return mk(-a.v, -a.d); }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return mk(a.v * b.v, a.v * b.d + a.d * b.v); }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
    double inv = 1.0 / b.v, q = a.v * inv;
    return mk(q, (a.d - q * b.d) * inv);
}
__device__ __forceinline__ Dual operator+(double a, Dual b) { return mk(a + b.v, b.d); }
__device__ __forceinline__ Dual operator-(double a, Dual b) { return mk(a - b.v, -b.d); }
__device__ __forceinline__ Dual operator*(double a, Dual b) { return mk(a * b.v, a * b.d); }
__device__ __forceinline__ Dual operator*(Dual a, double b) { return mk(a.v * b, a.d * b); }
__device__ __forceinline__ Dual sqrt_(Dual a) {
    double s = sqrt(a.v);
    return mk(s, a.d / (2.0 * s));
}

// ---------------------------------------------------------------------------------------
// second-order dual number  v + a e1 + b e2 + c e1 e2,  e1^2 = e2^2 = 0
// (d2f/dx_i dx_j from one evaluation with directions i and j)
// ---------------------------------------------------------------------------------------
struct Dual2 {
    double v, a, b, c;
};
__device__ __forceinline__ Dual2 mk2(double v, double a, double b, double c) {
    Dual2 r; r.v = v; r.a = a; r.b = b; r.c = c; return r;
}
__device__ __forceinline__ Dual2 operator+(Dual2 x, Dual2 y) { return mk2(x.v + y.v, x.a + y.a, x.b + y.b, x.c + y.c); }
__device__ __forceinline__ Dual2 operator-(Dual2 x, Dual2 y) { return mk2(x.v - y.v, x.a - y.a, x.b - y.b, x.c - y.c); }
__device__ __forceinline__ Dual2 operator-(Dual2 x) { return mk2(-x.v, -x.a, -x.b, -x.c); }
__device__ __forceinline__ Dual2 operator*(Dual2 x, Dual2 y) {
    return mk2(x.v * y.v, x.v * y.a + x.a * y.v, x.v * y.b + x.b * y.v,
               x.v * y.c + x.a * y.b + x.b * y.a + x.c * y.v);
}
__device__ __forceinline__ Dual2 recip_(Dual2 y) {
    double i = 1.0 / y.v, i2 = i * i;
    return mk2(i, -y.a * i2, -y.b * i2, (2.0 * y.a * y.b * i - y.c) * i2);
}
__device__ __forceinline__ Dual2 operator/(Dual2 x, Dual2 y) { return x * recip_(y); }
__device__ __forceinline__ Dual2 operator+(double a, Dual2 y) { return mk2(a + y.v, y.a, y.b, y.c); }
__device__ __forceinline__ Dual2 operator-(double a, Dual2 y) { return mk2(a - y.v, -y.a, -y.b, -y.c); }
__device__ __forceinline__ Dual2 operator*(double a, Dual2 y) { return mk2(a * y.v, a * y.a, a * y.b, a * y.c); }
__device__ __forceinline__ Dual2 operator*(Dual2 y, double a) { return a * y; }
__device__ __forceinline__ Dual2 sqrt_(Dual2 x) {
    double s = sqrt(x.v), h = 0.5 / s;
    return mk2(s, x.a * h, x.b * h, (x.c - 0.5 * x.a * x.b / x.v) * h);
}

__device__ __forceinline__ double sqrt_(double a) { return sqrt(a); }
// reciprocals: one division per scalar, the 3-vector scalings below are multiplies
__device__ __forceinline__ double rcp_(double a) { return 1.0 / a; }
__device__ __forceinline__ Dual rcp_(Dual a) { double i = 1.0 / a.v; return mk(i, -a.d * i * i); }
__device__ __forceinline__ Dual2 rcp_(Dual2 a) { return recip_(a); }
// x^(-1/2): norms and 1/sin come from one reciprocal square root instead of sqrt + divide
__device__ __forceinline__ double rsqrt_(double a) { return rsqrt(a); }
__device__ __forceinline__ Dual rsqrt_(Dual a) { double i = rsqrt(a.v); return mk(i, -0.5 * i * i * i * a.d); }
__device__ __forceinline__ Dual2 rsqrt_(Dual2 x) {
    double i = rsqrt(x.v), i2 = i * i, f1 = -0.5 * i * i2, f2 = 0.75 * i * i2 * i2;
    return mk2(i, f1 * x.a, f1 * x.b, f1 * x.c + f2 * x.a * x.b);
}
__device__ __forceinline__ double val(double a) { return a; }
__device__ __forceinline__ double val(Dual a) { return a.v; }
__device__ __forceinline__ double val(Dual2 a) { return a.v; }

// ---------------------------------------------------------------------------------------
// 3-vectors
// ---------------------------------------------------------------------------------------
template <class S> struct V3 {
    S x, y, z;
};
template <class S> __device__ __forceinline__ V3<S> v3(S x, S y, S z) { V3<S> r; r.x = x; r.y = y; r.z = z; return r; }
template <class S> __device__ __forceinline__ V3<S> operator+(V3<S> a, V3<S> b) { return v3<S>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class S> __device__ __forceinline__ V3<S> operator-(V3<S> a, V3<S> b) { return v3<S>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class S> __device__ __forceinline__ V3<S> operator-(V3<S> a) { return v3<S>(-a.x, -a.y, -a.z); }
template <class S> __device__ __forceinline__ V3<S> operator*(S s, V3<S> a) { return v3<S>(s * a.x, s * a.y, s * a.z); }
template <class S> __device__ __forceinline__ V3<S> operator/(V3<S> a, S s) { S i = rcp_(s); return v3<S>(a.x * i, a.y * i, a.z * i); }
template <class S> __device__ __forceinline__ S dot(V3<S> a, V3<S> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class S> __device__ __forceinline__ V3<S> cross(V3<S> a, V3<S> b) {
    return v3<S>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <class S> __device__ __forceinline__ S norm(V3<S> a) { return sqrt_(dot(a, a)); }
// |a| and 1/|a| together
template <class S> __device__ __forceinline__ void norm_inv(V3<S> a, S & n, S & inv) {
    S n2 = dot(a, a);
    inv = rsqrt_(n2);
    n = n2 * inv;
}

// ---------------------------------------------------------------------------------------
// angle post-processing shared by torsion / torsion2 (source/IC/InvDisp.cpp:108-118, 339-347)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double signed_wrapped_angle(double cosphi, double sign_test, double mn) {
    double phi;
    if (cosphi > 1.0) phi = 0.0;
    else if (cosphi < -1.0) phi = TC_PI;
    else {
        phi = acos(cosphi);
        if (sign_test < 0.0) phi = -phi;
    }
    if (phi < mn) phi += 2.0 * TC_PI;
    else if (phi > mn + 2.0 * TC_PI) phi -= 2.0 * TC_PI;
    return phi;
}

// ---------------------------------------------------------------------------------------
// J-path formulas (un-clamped), source/IC/InvDisp.cpp:256-634 (== :653-852).
// `a` are the atoms in the InvDisp's own order, `J` the per-atom gradient blocks, q the value
// as a scalar of type S (for the angle types the angle itself is returned separately as a
// double since only its value is ever needed).
// ---------------------------------------------------------------------------------------

// The reference divides vector by scalar component-wise; here every distinct denominator is
// inverted once and multiplied through (differs from the reference by rounding only).

// Every J function hands its per-atom blocks to `emit(i, block)` as soon as they exist (i is
// the atom's position in the InvDisp), so that no more than one block is live at a time.

// stretching :261-272
template <class S, class Emit> __device__ __forceinline__ void stretching_J(const V3<S> * a, S & q, Emit emit) {
    V3<S> r12 = a[1] - a[0];
    S n, i;
    norm_inv(r12, n, i);
    V3<S> u = i * r12;
    q = n;
    emit(0, -u);
    emit(1, u);
}

// bending :273-296 (angle = acos(q_cos)), cosbending :297-316
template <class S, bool COS, class Emit> __device__ __forceinline__ void bending_J(const V3<S> * a, S & costheta, Emit emit) {
    V3<S> r21 = a[0] - a[1];
    S i21 = rsqrt_(dot(r21, r21));
    V3<S> u21 = i21 * r21;
    V3<S> r23 = a[2] - a[1];
    S i23 = rsqrt_(dot(r23, r23));
    V3<S> u23 = i23 * r23;
    costheta = dot(u21, u23);
    V3<S> J0, J2;
    if (COS) {
        J0 = i21 * (u23 - costheta * u21);
        J2 = i23 * (u21 - costheta * u23);
    } else {
        S isin = rsqrt_(1.0 - costheta * costheta);
        J0 = (isin * i21) * (costheta * u21 - u23);
        J2 = (isin * i23) * (costheta * u23 - u21);
    }
    emit(0, J0);
    emit(2, J2);
    emit(1, -J0 - J2);
}

// torsion :317-354, sintorsion :355-384, costorsion :385-414
// VAR 0: torsion (J of the angle), 1: sintorsion, 2: costorsion.
// sinphi / cosphi / sign_test are returned for the caller to form q.
template <class S, int VAR, class Emit>
__device__ __forceinline__ void torsion_J(const V3<S> * a, S & sinphi, S & cosphi, S & sign_test, Emit emit) {
    V3<S> r12 = a[1] - a[0];
    S n12, i12;
    norm_inv(r12, n12, i12);
    V3<S> u12 = i12 * r12;
    V3<S> r23 = a[2] - a[1];
    S n23, i23;
    norm_inv(r23, n23, i23);
    V3<S> u23 = i23 * r23;
    V3<S> r34 = a[3] - a[2];
    S n34, i34;
    norm_inv(r34, n34, i34);
    V3<S> u34 = i34 * r34;
    S cos123 = -dot(u12, u23);
    S is123 = rsqrt_(1.0 - cos123 * cos123);
    S cos234 = -dot(u23, u34);
    S is234 = rsqrt_(1.0 - cos234 * cos234);
    V3<S> n123 = is123 * cross(u12, u23);
    V3<S> n234 = is234 * cross(u23, u34);
    cosphi = dot(n123, n234);
    if (VAR == 0) sign_test = dot(n123, cross(n234, u23));
    if (VAR != 0) sinphi = dot(cross(n123, n234), u23);
    // common factor of the sin / cos variants: d sin = cos dphi, d cos = -sin dphi
    S a23 = i23 * is123, b23 = i23 * is234;
    S w0 = -(i12 * is123), w1a = (n23 - n12 * cos123) * (i12 * a23), w1b = -(cos234 * b23);
    S w2a = (n34 * cos234 - n23) * (i34 * b23), w2b = cos123 * a23, w3 = i34 * is234;
    if (VAR == 1) { w0 = cosphi * w0; w1a = cosphi * w1a; w1b = cosphi * w1b; w2a = cosphi * w2a; w2b = cosphi * w2b; w3 = cosphi * w3; }
    if (VAR == 2) { S m = -sinphi; w0 = m * w0; w1a = m * w1a; w1b = m * w1b; w2a = m * w2a; w2b = m * w2b; w3 = m * w3; }
    emit(0, w0 * n123);
    emit(1, w1a * n123 + w1b * n234);
    emit(2, w2a * n234 + w2b * n123);
    emit(3, w3 * n234);
}

// OutOfPlane :571-602 (angle = asin(sintheta)), sinoop :603-631
template <class S, bool SIN, class Emit> __device__ __forceinline__ void oop_J(const V3<S> * a, S & sintheta, Emit emit) {
    V3<S> r21 = a[0] - a[1];
    S i21 = rsqrt_(dot(r21, r21));
    V3<S> u21 = i21 * r21;
    V3<S> r23 = a[2] - a[1];
    S i23 = rsqrt_(dot(r23, r23));
    V3<S> u23 = i23 * r23;
    V3<S> r24 = a[3] - a[1];
    S i24 = rsqrt_(dot(r24, r24));
    V3<S> u24 = i24 * r24;
    S cos324 = dot(u23, u24);
    S sin324sq = 1.0 - cos324 * cos324;
    S is324 = rsqrt_(sin324sq);
    S is324sq = is324 * is324;
    V3<S> c41 = cross(u24, u21);
    sintheta = dot(u23, c41) * is324;
    S w, t, s1;            // J0 = i21 (w u23 x u24 - s1 u21), J2 = i23 (w u24 x u21 - t (u23 - c u24)), ...
    if (SIN) {
        w = is324; t = sintheta * is324sq; s1 = sintheta;
    } else {
        S icos = rsqrt_(1.0 - sintheta * sintheta);
        S tantheta = sintheta * icos;
        w = icos * is324; t = tantheta * is324sq; s1 = tantheta;
    }
    V3<S> J0 = i21 * (w * cross(u23, u24) - s1 * u21);
    emit(0, J0);
    V3<S> acc = J0;
    V3<S> J2 = i23 * (w * c41 - t * (u23 - cos324 * u24));
    emit(2, J2);
    acc = acc + J2;
    V3<S> J3 = i24 * (w * cross(u21, u23) - t * (u24 - cos324 * u23));
    emit(3, J3);
    emit(1, -(acc + J3));
}

// sin / cos of the 5-atom dihedral, source/IC/InvDisp.cpp:419-429 (value expression that the
// reference hands to autograd)
template <class S>
__device__ __forceinline__ void torsion2_sincos(const V3<S> * a, S & sinphi, S & cosphi, S & sign_test) {
    V3<S> r12 = a[1] - a[0];
    V3<S> r23 = a[2] - a[1];
    V3<S> r45 = a[4] - a[3];
    V3<S> n123 = cross(r12, r23);
    n123 = n123 / norm(n123);
    V3<S> n2345 = cross(r23, r45);
    n2345 = n2345 / norm(n2345);
    sinphi = dot(cross(n123, n2345), r23 / norm(r23));
    cosphi = dot(n123, n2345);
    sign_test = dot(n123, cross(n2345, r23));
}

// ---------------------------------------------------------------------------------------
// value path, the clamped formulas of InvDisp::operator() (source/IC/InvDisp.cpp:64-254)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double value_path(int type, const V3<double> * a, double mn) {
    switch (type) {
    case 0: return 1.0;                                                    // dummy :68
    case 1: return norm(a[1] - a[0]);                                      // stretching :69-73
    case 2: case 3: {                                                      // bending :74-88, cosbending :89-97
        V3<double> r21 = a[0] - a[1]; r21 = r21 / norm(r21);
        V3<double> r23 = a[2] - a[1]; r23 = r23 / norm(r23);
        double c = dot(r21, r23);
        if (type == 3) return c;
        if (c > 1.0) return 0.0;
        if (c < -1.0) return TC_PI;
        return acos(c);
    }
    case 4: case 5: case 6: {                                              // torsion family :98-142
        V3<double> r12 = a[1] - a[0], r23 = a[2] - a[1], r34 = a[3] - a[2];
        V3<double> n123 = cross(r12, r23); n123 = n123 / norm(n123);
        V3<double> n234 = cross(r23, r34); n234 = n234 / norm(n234);
        if (type == 5) return dot(cross(n123, n234), r23 / norm(r23));
        double cosphi = dot(n123, n234);
        if (type == 6) return cosphi;
        return signed_wrapped_angle(cosphi, dot(n123, cross(n234, r23)), mn);
    }
    case 7: case 8: case 9: case 10: case 11: {                            // torsion2 family :143-225
        V3<double> r12 = a[1] - a[0], r23 = a[2] - a[1], r45 = a[4] - a[3];
        V3<double> n123 = cross(r12, r23); n123 = n123 / norm(n123);
        V3<double> n2345 = cross(r23, r45); n2345 = n2345 / norm(n2345);
        double cosphi = dot(n123, n2345);
        if (type == 9) return cosphi;
        if (type == 7) return signed_wrapped_angle(cosphi, dot(n123, cross(n2345, r23)), mn);
        V3<double> u23 = r23 / norm(r23);
        double sinphi = dot(cross(n123, n2345), u23);
        if (type == 8) return sinphi;
        V3<double> u12 = r12 / norm(r12);
        double costheta = -dot(u12, u23);
        double sintheta = sqrt(1.0 - costheta * costheta);
        return sintheta * (type == 10 ? cosphi : sinphi);
    }
    default: {                                                             // OutOfPlane :226-241, sinoop :242-251
        V3<double> r21 = a[0] - a[1], r23 = a[2] - a[1], r24 = a[3] - a[1];
        V3<double> n234 = cross(r23, r24); n234 = n234 / norm(n234);
        double s = dot(n234, r21 / norm(r21));
        if (type == 13) return s;
        if (s < -1.0) return -TC_PI / 2.0;
        if (s > 1.0) return TC_PI / 2.0;
        return asin(s);
    }
    }
}

} // namespace tchem_b200
