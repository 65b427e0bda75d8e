// metric.cuh -- bit-exact edge-cost functors for sm_100a.
//
// The reference's ONLY cost routine is src/tsp.c:629
//     (double)((int)(sqrtf(pow(dx,2) + pow(dy,2)) + 0.5))
// which gcc -O3 (no -ffast-math, baseline x86-64 => no FMA) turns into
//     subsd subsd mulsd mulsd addsd cvtsd2ss sqrtss cvtss2sd addsd(0.5) cvttsd2si cvtsi2sd.
// EucGeneral / EucInt reproduce that value for every input whose cost is below 2^23 (checked at
// context creation) with 4-5 FP64 instructions, one F2F.F32.F64, one MUFU.RSQ and 6-7 FP32 ones (scalar
// form K) or 3 packed FP32x2 ones per cost (scan form K2, two costs at a time):
//
//   * dx, dy, dx*dx, dy*dy, sum: explicit round-to-nearest intrinsics, so nvcc cannot contract
//     dx*dx + dy*dy into an FMA (with non-integer coordinates such as usa13509's that would
//     change the double before the float rounding): EucGeneral.  With all-integer coordinates
//     below 2^25 every intermediate is an exact integer, so EucInt may use FMAs (4 FP64
//     instructions) and folds the zero guard into them (see there).
//   * sqrtf: sqrt_rn_pos() is the IEEE round-to-nearest square root for positive normal floats
//     (MUFU.RSQ seed + one Newton step with an exact FMA residual).  It is checked exhaustively
//     against sqrt.rn.f32 on the device by tspcuda_selfcheck_sqrt (tests/test_gpu_parity.py).
//   * (int)(r + 0.5) for r >= 0 is floor(r + 0.5).  r + 0.5 is formed in FP32 with round-DOWN
//     (v <= r+0.5 and v >= floor(r+0.5) because that integer is representable => same floor), and
//     adding 2^23 with round-down leaves 2^23 + floor(v) in the float, whose bit pattern is
//     0x4B000000 + cost.  Costs therefore travel as "K-offset" int32 values; the offset cancels
//     in every 2-opt delta (two costs added, two subtracted), so no F2I/I2F conversion is needed.
//
// CEIL_2D and ATT follow TSPLIB95 in FP64 (IEEE sqrt/div), "parity unpinned": the reference
// rejects those files (src/tsp.c:576-583).
#pragma once
#include <cuda_runtime.h>

namespace tspcuda {

constexpr int KOFF = 0x4B000000;        // bit pattern of 8388608.0f (2^23)
constexpr int BIGPEN = 1 << 27;         // penalty that removes a pair from the argmin
constexpr int DELTA_BIAS = 1 << 29;     // bias of the delta field in the packed key
constexpr unsigned long long KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;
#define TINY_F 8.673617379884035e-19f        /* 2^-60: keeps rsqrt finite; any f < 0.25 costs 0 */
#define TINY_D 8.673617379884035e-19

// IEEE-754 round-to-nearest sqrt for f in [2^-60, 2^64] (normal, positive).
__device__ __forceinline__ float sqrt_rn_pos(float f) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(f));
    float g = __fmul_rn(f, y);
    float h = __fmul_rn(y, 0.5f);
    float e = __fmaf_rn(-g, g, f);
    return __fmaf_rn(e, h, g);
}

// K-offset cost from the float that the reference hands to sqrtf().
__device__ __forceinline__ int koff_from_sumsq(float f) {     // f > 0 (callers clamp or add 2^-60)
    float r = sqrt_rn_pos(f);
    float v = __fadd_rd(r, 0.5f);
    float w = __fadd_rd(v, 8388608.0f);
    return __float_as_int(w);
}

// ---- the same for TWO costs at once with Blackwell's packed FP32x2 instructions (FMUL2 / FFMA2 / FADD2):
// half the FP32 issue slots of the scan.  Inputs are the NEGATED sums of squares (the FP64 chain produces
// -s for free through operand negation), which removes every other negation: with g' = nf*y = -g,
//   e' = g'*g' + nf = g^2 - f,   r' = h*e' + g' = -(g + h*(f - g^2)) = -r,
// and floor(r + 0.5) + 2^23 rounded DOWN becomes the same two additions of -0.5 and -2^23 rounded UP
// (RM(x) = -RP(-x)).  The results therefore carry the bit pattern 0xCB000000 + cost, i.e. the K-offset
// encoding with the sign bit set: SCAN_FLIP.  Inside a scan every delta adds two such costs and subtracts
// two K-offset tour-edge costs, so the extra 2 * 2^31 wraps away in 32-bit arithmetic.  Checked
// exhaustively against koff_from_sumsq on the device (tspcuda_selfcheck_sqrt).
constexpr unsigned SCAN_FLIP = 0x80000000u;

__device__ __forceinline__ void koff2_from_negsumsq(float nf0, float nf1, int &k0, int &k1) {
    float y0, y1;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(fabsf(nf0)));
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y1) : "f"(fabsf(nf1)));
    unsigned long long nf, y, g, h, e, r, v, w;
    asm("mov.b64 %0, {%1, %2};" : "=l"(nf) : "f"(nf0), "f"(nf1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(y0), "f"(y1));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(g) : "l"(nf), "l"(y));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(h) : "l"(y), "l"(0x3F0000003F000000ull));
    asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(e) : "l"(g), "l"(nf));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(h), "l"(e), "l"(g));
    asm("add.rp.f32x2 %0, %1, %2;" : "=l"(v) : "l"(r), "l"(0xBF000000BF000000ull));
    asm("add.rp.f32x2 %0, %1, %2;" : "=l"(w) : "l"(v), "l"(0xCB000000CB000000ull));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(k0), "=r"(k1) : "l"(w));
}

// Scan-side pair interface of every metric: K2 = the costs of (row a, column) and (row b, column) in
// the metric's scan encoding K-offset ^ SCAN_XOR (sweep_block, kernels.cuh).
template <class Derived>
struct ScalarScan {
    static constexpr unsigned SCAN_XOR = 0;
    static __device__ __forceinline__ void K2(double xa, double ya, double xb, double yb, double xj, double yj, int &ka, int &kb) {
        ka = Derived::K(xa, ya, xj, yj);
        kb = Derived::K(xb, yb, xj, yj);
    }
    // two unrelated edges (p,q) and (r,s) at once, same encoding (patching.cuh)
    static __device__ __forceinline__ void KK(double xp, double yp, double xq, double yq, double xr, double yr, double xs, double ys,
                                              int &ka, int &kb) {
        ka = Derived::K(xp, yp, xq, yq);
        kb = Derived::K(xr, yr, xs, ys);
    }
};

// tsp.c:629 for arbitrary (non-integer) coordinates: separate roundings, clamp against s == 0.
struct EucGeneral {
    static __device__ __forceinline__ int K(double xi, double yi, double xj, double yj) {
        double dx = __dsub_rn(xj, xi);
        double dy = __dsub_rn(yj, yi);
        double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        return koff_from_sumsq(fmaxf(__double2float_rn(s), TINY_F));
    }
    static constexpr unsigned SCAN_XOR = SCAN_FLIP;
    static __device__ __forceinline__ void K2(double xa, double ya, double xb, double yb, double xj, double yj, int &ka, int &kb) {
        const double dxa = __dsub_rn(xj, xa), dya = __dsub_rn(yj, ya), dxb = __dsub_rn(xj, xb), dyb = __dsub_rn(yj, yb);
        const double sa = __dadd_rn(__dmul_rn(dxa, dxa), __dmul_rn(dya, dya));
        const double sb = __dadd_rn(__dmul_rn(dxb, dxb), __dmul_rn(dyb, dyb));
        koff2_from_negsumsq(fminf(__double2float_rn(-sa), -TINY_F), fminf(__double2float_rn(-sb), -TINY_F), ka, kb);
    }
    static __device__ __forceinline__ void KK(double xp, double yp, double xq, double yq, double xr, double yr, double xs, double ys,
                                              int &ka, int &kb) {
        const double dxa = __dsub_rn(xq, xp), dya = __dsub_rn(yq, yp), dxb = __dsub_rn(xs, xr), dyb = __dsub_rn(ys, yr);
        const double sa = __dadd_rn(__dmul_rn(dxa, dxa), __dmul_rn(dya, dya));
        const double sb = __dadd_rn(__dmul_rn(dxb, dxb), __dmul_rn(dyb, dyb));
        koff2_from_negsumsq(fminf(__double2float_rn(-sa), -TINY_F), fminf(__double2float_rn(-sb), -TINY_F), ka, kb);
    }
    // Euclidean radius that contains every w with cost(u,w) < r (pruned.cuh): cost <= r-1 means
    // sqrtf(f) < r - 0.5 with f, sqrtf within 2^-23 relative of the exact values.
    static __device__ __forceinline__ double radius(double r) { return r * 1.000001 + 1.0; }
    // a cost every node at Euclidean distance >= d is guaranteed to reach (grid-accelerated NN, pruned.cuh)
    static __device__ __forceinline__ double cost_floor(double d) { return d * 0.999999 - 1.0; }
};

// All-integer coordinates below 2^25: every intermediate is an exact integer < 2^53, so FMAs are
// allowed.  The 2^-60 addend only survives when dx == dy == 0 (coincident nodes, cost 0) and keeps
// rsqrt finite there without a clamp instruction.
struct EucInt {
    static __device__ __forceinline__ int K(double xi, double yi, double xj, double yj) {
        double dx = __dsub_rn(xj, xi);
        double dy = __dsub_rn(yj, yi);
        double s = __fma_rn(dx, dx, __fma_rn(dy, dy, TINY_D));
        return koff_from_sumsq(__double2float_rn(s));
    }
    static constexpr unsigned SCAN_XOR = SCAN_FLIP;
    static __device__ __forceinline__ void K2(double xa, double ya, double xb, double yb, double xj, double yj, int &ka, int &kb) {
        const double dxa = __dsub_rn(xj, xa), dya = __dsub_rn(yj, ya), dxb = __dsub_rn(xj, xb), dyb = __dsub_rn(yj, yb);
        const double nsa = __fma_rn(-dxa, dxa, __fma_rn(-dya, dya, -TINY_D));      // -(dx^2 + dy^2 + 2^-60), exact
        const double nsb = __fma_rn(-dxb, dxb, __fma_rn(-dyb, dyb, -TINY_D));
        koff2_from_negsumsq(__double2float_rn(nsa), __double2float_rn(nsb), ka, kb);
    }
    static __device__ __forceinline__ void KK(double xp, double yp, double xq, double yq, double xr, double yr, double xs, double ys,
                                              int &ka, int &kb) {
        const double dxa = __dsub_rn(xq, xp), dya = __dsub_rn(yq, yp), dxb = __dsub_rn(xs, xr), dyb = __dsub_rn(ys, yr);
        const double nsa = __fma_rn(-dxa, dxa, __fma_rn(-dya, dya, -TINY_D));
        const double nsb = __fma_rn(-dxb, dxb, __fma_rn(-dyb, dyb, -TINY_D));
        koff2_from_negsumsq(__double2float_rn(nsa), __double2float_rn(nsb), ka, kb);
    }
    static __device__ __forceinline__ double radius(double r) { return r * 1.000001 + 1.0; }
    static __device__ __forceinline__ double cost_floor(double d) { return d * 0.999999 - 1.0; }
};

// MEASUREMENT ONLY (tspcuda_pipe_rate kind 10, profiles/r02_k2.md): the same two costs for integer coordinates below 2^23
// WITHOUT any FP64 instruction and without F2F.F32.F64 -- the float the reference hands to sqrtf(), RN32(dx^2 + dy^2),
// from an error-free transformation in FP32: dx, dy exact; q = -(dx*dx) rounded, e = fma(dx, dx, q) its exact residual;
// TwoSum of the two rounded squares; the residuals are integers below 2^24, so their sum is exact and ONE final rounding
// add gives the correctly rounded (negated) sum.  15 packed FP32x2 instructions + 4 LOP3 + 2 FMNMX per two costs replace
// 4 DADD + 4 DFMA + 2 F2F: one XU operation per cost instead of two, but 21 packed FP32 instructions per two costs on a
// pipe that retires one packed instruction per two cycles.  It measures SLOWER than EucInt::K2 (see the profile note), so
// no kernel uses it.
struct EucIntF32 {
    static __device__ __forceinline__ void K2(float xa, float ya, float xb, float yb, float xj, float yj, int &ka, int &kb) {
        unsigned long long X, Y, XJ, YJ, dx, dy, ndx, ndy, q1, q2, e1, e2, sm, bb, t0, t1, t, es, rn, nf;
        asm("mov.b64 %0, {%1, %2};" : "=l"(X) : "f"(xa), "f"(xb));
        asm("mov.b64 %0, {%1, %2};" : "=l"(Y) : "f"(ya), "f"(yb));
        asm("mov.b64 %0, {%1, %1};" : "=l"(XJ) : "f"(xj));
        asm("mov.b64 %0, {%1, %1};" : "=l"(YJ) : "f"(yj));
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(XJ), "l"(X));
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(YJ), "l"(Y));
        ndx = dx ^ 0x8000000080000000ull;
        ndy = dy ^ 0x8000000080000000ull;
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(q1) : "l"(dx), "l"(ndx));               // -(dx^2), rounded
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(q2) : "l"(dy), "l"(ndy));
        asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(e1) : "l"(dx), "l"(q1));            // dx^2 + q1: exact
        asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(e2) : "l"(dy), "l"(q2));
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(sm) : "l"(q1), "l"(q2));                // TwoSum(q1, q2) = (sm, t)
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(bb) : "l"(sm), "l"(q1));
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(t0) : "l"(sm), "l"(bb));
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(t0) : "l"(q1), "l"(t0));
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(t1) : "l"(q2), "l"(bb));
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(t0), "l"(t1));
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(es) : "l"(e1), "l"(e2));
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rn) : "l"(t), "l"(es));                 // exact: integers below 2^24
        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(nf) : "l"(sm), "l"(rn));                // -RN32(dx^2 + dy^2)
        float n0, n1;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(n0), "=f"(n1) : "l"(nf));
        koff2_from_negsumsq(fminf(n0, -TINY_F), fminf(n1, -TINY_F), ka, kb);
    }
};

// TSPLIB95 CEIL_2D: ceil(sqrt(dx^2+dy^2)) in double.
struct Ceil2D : ScalarScan<Ceil2D> {
    static __device__ __forceinline__ int K(double xi, double yi, double xj, double yj) {
        double dx = __dsub_rn(xj, xi);
        double dy = __dsub_rn(yj, yi);
        double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        return KOFF + (int)ceil(__dsqrt_rn(s));
    }
    static __device__ __forceinline__ double radius(double r) { return r * 1.000001 + 1.0; }      // ceil(d) < r => d < r
    static __device__ __forceinline__ double cost_floor(double d) { return d * 0.999999 - 1.0; }
};

// TSPLIB95 ATT: r = sqrt((dx^2+dy^2)/10); t = nint(r); t < r ? t+1 : t.
struct Att : ScalarScan<Att> {
    static __device__ __forceinline__ int K(double xi, double yi, double xj, double yj) {
        double dx = __dsub_rn(xj, xi);
        double dy = __dsub_rn(yj, yi);
        double s = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        double r = __dsqrt_rn(__ddiv_rn(s, 10.0));
        double t = (double)(long long)__dadd_rn(r, 0.5);
        return KOFF + (int)((t < r) ? t + 1.0 : t);
    }
    // cost >= sqrt(s/10) - 1  =>  cost < r implies d < (r + 1) * sqrt(10)
    static __device__ __forceinline__ double radius(double r) { return (r + 1.0) * 3.1622777 + 1.0; }
    static __device__ __forceinline__ double cost_floor(double d) { return d * 0.3162277 - 2.0; }
};

// TSPLIB95 GEO (section 2.4).  The node "coordinates" this functor sees are (latitude, longitude) in
// radians, converted once per instance by geo_convert_kernel with the TSPLIB constants (PI = 3.141592,
// DDD.MM input).  Same operation order as the definition, no FMA contraction; cos/acos are CUDA's
// double-precision routines (<= 2 ulp), so a cost can differ from a libm evaluation only where
// RRR*acos(.)+1 lies within ~1e-8 of an integer.  Parity unpinned (the reference has no GEO rule).
struct Geo : ScalarScan<Geo> {
    static __device__ __forceinline__ int K(double lati, double loni, double latj, double lonj) {
        const double q1 = cos(__dsub_rn(loni, lonj));
        const double q2 = cos(__dsub_rn(lati, latj));
        const double q3 = cos(__dadd_rn(lati, latj));
        const double f = __dmul_rn(0.5, __dsub_rn(__dmul_rn(__dadd_rn(1.0, q1), q2), __dmul_rn(__dsub_rn(1.0, q1), q3)));
        return KOFF + (int)__dadd_rn(__dmul_rn(6378.388, acos(f)), 1.0);
    }
    static __device__ __forceinline__ double radius(double r) { return r; }     // unused: no pruned scan for GEO
    static __device__ __forceinline__ double cost_floor(double) { return -1.0; }  // unused
};

__device__ __forceinline__ double geo_radians(double v) {
    const double deg = (double)(long long)v;                  // truncation toward zero
    const double min = __dsub_rn(v, deg);
    return __ddiv_rn(__dmul_rn(3.141592, __dadd_rn(deg, __ddiv_rn(__dmul_rn(5.0, min), 3.0))), 180.0);
}

__device__ __forceinline__ unsigned long long pack_key(int delta, int u, int v) {
    unsigned lo = (unsigned)min(u, v), hi = (unsigned)max(u, v);
    return ((unsigned long long)(unsigned)(delta + DELTA_BIAS) << 34) | ((unsigned long long)lo << 17) |
           (unsigned long long)hi;
}
__device__ __host__ __forceinline__ int key_delta(unsigned long long k) { return (int)(k >> 34) - DELTA_BIAS; }
__device__ __host__ __forceinline__ int key_a(unsigned long long k) { return (int)((k >> 17) & 0x1FFFF); }
__device__ __host__ __forceinline__ int key_b(unsigned long long k) { return (int)(k & 0x1FFFF); }

}  // namespace tspcuda
