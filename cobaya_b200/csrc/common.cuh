// common.cuh -- device-side model description, Philox4x32-10 streams and the draw
// primitives shared by every kernel of the B200 ensemble-MCMC engine (sm_100a).
//
// Randomness replaces numpy's Generator calls on the reference hot path
// (cobaya/samplers/mcmc/proposal.py:54,79-82,90; cobaya/functions.py:36;
// cobaya/samplers/mcmc/mcmc.py:683) with counter-based streams, so that any
// thread can produce any draw and the CPU oracle consumes identical values
// (layout: DESIGN.md section 4).
#pragma once
#ifdef __CUDACC_RTC__
// run-time compilation (ext_functor.inl: the general step kernel with the user's functions
// inlined): NVRTC has no host headers
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#define CUDART_INF (__longlong_as_double(0x7ff0000000000000LL))
#define CUDART_NAN (__longlong_as_double(0xfff8000000000000LL))
#define CUDART_PI 3.1415926535897931e+0
#include "cobaya_b200.h"
#else
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "../../include/cobaya_b200.h"
#endif

#define CB2_TAG_STEP 0u
#define CB2_TAG_ACCEPT 1u
#define CB2_TAG_BASIS 2u
#define CB2_TAG_CYCLER 3u
#define CB2_TAG_STEP2 4u

#define CB2_LOG_2PI 1.8378770664093453

// Shape part f(z; a, b) of the 1-D prior log-densities the reference takes from
// scipy.stats (cobaya/prior.py:520-525, cobaya/tools.py:611-718):
//   logpdf(x) = cn + f((x - loc) / scale; a, b),  cn precomputed on the host.
// Kinds: include/cobaya_b200.h (CB2_PRIOR_*).  Called only inside the support
// (Prior.logps_internal checks the bounds first, prior.py:752).
__device__ __forceinline__ double prior1d_shape(int kind, double z, double a, double b) {
    switch (kind) {
        case CB2_PRIOR_NORMAL:
        case CB2_PRIOR_TRUNCNORM:
        case CB2_PRIOR_HALFNORM: return -z * z / 2;
        case CB2_PRIOR_EXPON: return -z;
        case CB2_PRIOR_BETA: {
            double s = 0.0;
            if (a != 1.0) s += (a - 1.0) * log(z);
            if (b != 1.0) s += (b - 1.0) * log1p(-z);
            return s;
        }
        case CB2_PRIOR_GAMMA: return (a != 1.0 ? (a - 1.0) * log(z) : 0.0) - z;
        case CB2_PRIOR_LOGNORM: {
            if (!(z > 0.0)) return -CUDART_INF;
            const double l = log(z);
            return -l - l * l / (2 * a * a);
        }
        case CB2_PRIOR_CAUCHY: return -log1p(z * z);
        case CB2_PRIOR_LAPLACE: return -fabs(z);
        case CB2_PRIOR_LOGUNIFORM: return (z > 0.0) ? -log(z) : -CUDART_INF;
        default: return 0.0;
    }
}

struct LikeDev {
    int32_t kind, dim, n_modes, derived;
    int32_t idx_off;     // into Model.ipool
    int32_t means_off;   // into Model.dpool: [m*dim]
    int32_t linvT_off;   // [m][dim*dim], element (row i, col j) at j*dim + i
    int32_t c0_off;      // [m]  dim*log(2pi) + logdet_k
    int32_t w_off;       // [m]
    int32_t der_off;     // offset of this likelihood's derived params in the row
    double scale;        // rosenbrock
};

struct ModelDev {
    int32_t D, n_like, n_der, width;
    int32_t n_ep;   // external priors: extra minuslogprior__<name> columns in a row
    // prior
    const int32_t *prior_kind;  // [D]
    const double *lower, *upper, *loc, *pscale;
    const int32_t *periodic;
    // shape parameters and normalisation of the non-uniform, non-normal 1-D priors
    // (cb2_set_prior_shapes); null when every parameter is uniform or normal
    const double *pa, *pb, *pcn;
    // any_normal: some parameter has a non-uniform prior; any_generic: some kind >= 2
    int32_t any_periodic, any_normal, any_generic;
    double uniform_logp;
    // likelihoods
    LikeDev likes[CB2_MAX_LIKES];
    const double *dpool;
    const int32_t *ipool;
    // blocking
    int32_t n_blocks;
    int32_t bsize[CB2_MAX_BLOCKS], jstart[CB2_MAX_BLOCKS], oversamp[CB2_MAX_BLOCKS];
    const int32_t *i_of_j;  // [D]
    int32_t drag, last_slow, drag_steps, n_slow, n_fast;
    // proposal: TT[k*D + j] = T[j][k]  (column k contiguous over rows j)
    const double *TT;
    double proposal_scale;
    // options
    double temperature;
    int64_t max_tries;
    int32_t output_thin;
    // rng
    uint32_t key0, key1;
    uint64_t chain_id0;
};

// ---------------------------------------------------------------- Philox4x32-10
struct u32x4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ u32x4 philox4x32_10(uint32_t k0, uint32_t k1,
                                                         uint32_t c0, uint32_t c1,
                                                         uint32_t c2, uint32_t c3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    u32x4 o;
    o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// 52-bit uniform strictly inside (0,1): (k + 1/2) 2^-52 with k = 26+26 bits of two words.
// On the device it is assembled from the bit pattern of 1 + k 2^-52 (no 64-bit int->double
// conversion, which runs on the slow XU pipe); both forms are exact and bit-identical.
__host__ __device__ __forceinline__ double u52(uint32_t a, uint32_t b) {
    uint64_t k = ((uint64_t)(a >> 6) << 26) | (uint64_t)(b >> 6);
#ifdef __CUDA_ARCH__
    return (__longlong_as_double((long long)(0x3FF0000000000000ull | k)) - 1.0) +
           1.1102230246251565e-16;  // 2^-53
#else
    return ((double)k + 0.5) * 2.220446049250313e-16;
#endif
}

// (w + 1/2) 2^-32 for a 32-bit word, same construction
__host__ __device__ __forceinline__ double u32half(uint32_t w) {
#ifdef __CUDA_ARCH__
    return (__longlong_as_double((long long)(0x4330000000000000ull | (uint64_t)w)) -
            4503599627370496.0 + 0.5) * 2.3283064365386963e-10;
#else
    return ((double)w + 0.5) * 2.3283064365386963e-10;
#endif
}

// radial part of a proposal: propose_r / RandProposer1D (proposal.py:71-93)
__device__ __forceinline__ void draw_radial(const ModelDev &M, uint64_t gid, uint64_t t,
                                            uint32_t sub, int n_block, double &r,
                                            double &sign) {
    u32x4 w = philox4x32_10(M.key0, M.key1, (uint32_t)t, (uint32_t)(t >> 32),
                            (uint32_t)gid, CB2_TAG_STEP | (sub << 8));
    double u_mix = u32half(w.x);
    double u_r = u52(w.z, w.w);
    sign = (w.y & 1u) ? 1.0 : -1.0;
    if (u_mix < 0.33) {
        r = -log(u_r);
    } else if (n_block >= 2) {
        r = sqrt(-2.0 * log(u_r));
    } else {
        u32x4 w2 = philox4x32_10(M.key0, M.key1, (uint32_t)t, (uint32_t)(t >> 32),
                                 (uint32_t)gid, CB2_TAG_STEP2 | (sub << 8));
        double u2 = u52(w2.x, w2.y);
        r = fabs(sqrt(-2.0 * log(u_r)) * cospi(2.0 * u2));
    }
}

// Exp(1) draw of metropolis_accept (mcmc.py:683)
__device__ __forceinline__ double draw_accept_exp(const ModelDev &M, uint64_t gid,
                                                  uint64_t t, uint32_t sub) {
    u32x4 w = philox4x32_10(M.key0, M.key1, (uint32_t)t, (uint32_t)(t >> 32),
                            (uint32_t)gid, CB2_TAG_ACCEPT | (sub << 8));
    return -log(u52(w.x, w.y));
}

// sin(2 pi u), cos(2 pi u) for u = (k + 1/2) 2^-52 given the 52-bit integer k.  The octant
// comes straight from the top bits of k and the reduced argument f = 2u - q/2 in
// [-1/4, 1/4] is assembled exactly from the remaining bits (no FRND/F2I: those run on the
// slow XU pipe); sin(pi f), cos(pi f) by Taylor series (|pi f| <= 0.786, error < 1 ulp).
__device__ __forceinline__ void sincos2pi_from_bits(uint64_t k, double &s, double &c) {
    const int q = (int)((k + (1ull << 49)) >> 50);                 // round(4u), 0..4
    const long long kk = (long long)k - ((long long)q << 50);      // in [-2^49, 2^49)
    const double kd = __longlong_as_double((long long)(0x4330000000000000ull |
                                                       (uint64_t)(kk + (1ll << 51)))) -
                      6755399441055744.0;                          // 2^52 + 2^51
    const double f = (kd + 0.5) * 4.440892098500626e-16;           // 2^-51
    const double x = 3.141592653589793 * f, x2 = x * x;
    double ps = -7.647163731819816e-13;                            // -1/15!
    ps = fma(ps, x2, 1.6059043836821613e-10);                      //  1/13!
    ps = fma(ps, x2, -2.505210838544172e-08);                      // -1/11!
    ps = fma(ps, x2, 2.7557319223985893e-06);                      //  1/9!
    ps = fma(ps, x2, -1.984126984126984e-04);                      // -1/7!
    ps = fma(ps, x2, 8.333333333333333e-03);                       //  1/5!
    ps = fma(ps, x2, -1.6666666666666666e-01);                     // -1/3!
    const double sf = fma(x * x2, ps, x);
    double pc = 4.779477332387385e-14;                             //  1/16!
    pc = fma(pc, x2, -1.1470745597729725e-11);                     // -1/14!
    pc = fma(pc, x2, 2.08767569878681e-09);                        //  1/12!
    pc = fma(pc, x2, -2.755731922398589e-07);                      // -1/10!
    pc = fma(pc, x2, 2.48015873015873e-05);                        //  1/8!
    pc = fma(pc, x2, -1.388888888888889e-03);                      // -1/6!
    pc = fma(pc, x2, 4.1666666666666664e-02);                      //  1/4!
    pc = fma(pc, x2, -0.5);
    const double cf = fma(pc, x2, 1.0);
    switch (q & 3) {
        case 0: s = sf; c = cf; break;
        case 1: s = cf; c = -sf; break;
        case 2: s = -sf; c = -cf; break;
        default: s = -cf; c = sf; break;
    }
}

// ---- table-driven natural logarithm for the Box-Muller radius of the basis kernels ----
// u = m 2^e with m in [0.75, 1.5); c_j = 0.75 + j/128 the table point nearest to m
// (j = 0..96, c_32 = 1 exactly); log u = e ln2 + log c_j + log1p((m - c_j)/c_j), the last
// term a degree-7 series in |r| <= 1/192 (truncation < 1e-19).  m - c_j is exact, so the
// result keeps full relative accuracy near u = 1; max error about 1 ulp.  Everything stays
// on the FMA pipe: exponent and index come from the bit pattern (no I2F/F2I).
#define CB2_LOGTAB_N 97
#define CB2_LOGTAB_DOUBLES (3 * CB2_LOGTAB_N)
__device__ double g_logtab[CB2_LOGTAB_DOUBLES];  // [j] = {c_j, 1/c_j, log c_j}; cb2_create fills it

__device__ __forceinline__ double log_tab(double u, const double *tab) {
    const uint64_t b = (uint64_t)__double_as_longlong(u);
    const uint64_t mant = b & 0x000FFFFFFFFFFFFFull;
    const int hi = (int)(mant >> 51);                         // mantissa >= 1.5: use m/2, e+1
    const int ex = (int)(b >> 52) - 1023 + hi;
    const double m = __longlong_as_double((long long)(mant | ((uint64_t)(0x3FF - hi) << 52)));
    const int j = hi ? (int)((mant + (1ull << 45)) >> 46) - 32
                     : 32 + (int)((mant + (1ull << 44)) >> 45);
    const double *t = tab + 3 * j;
    const double r = (m - t[0]) * t[1];
    double p = 1.0 / 7.0;
    p = fma(p, r, -1.0 / 6.0);
    p = fma(p, r, 0.2);
    p = fma(p, r, -0.25);
    p = fma(p, r, 1.0 / 3.0);
    p = fma(p, r, -0.5);
    const double l1p = fma(r * r, p, r);
    const double ed = __longlong_as_double((long long)(0x4330000000000000ull +
                                                       (uint64_t)(ex + 2048))) -
                      (4503599627370496.0 + 2048.0);
    return fma(ed, 0.6931471805599453, t[2] + l1p);
}

// pair p of the standard normals consumed by random_SO_N (functions.py:36)
__device__ __forceinline__ void draw_normal_pair(uint32_t k0, uint32_t k1, uint64_t gid,
                                                 int block, uint32_t epoch, uint32_t p,
                                                 double &z0, double &z1) {
    u32x4 w = philox4x32_10(k0, k1, p, epoch, (uint32_t)gid,
                            CB2_TAG_BASIS | ((uint32_t)block << 8));
    const double u1 = u52(w.x, w.y);
    const uint64_t k2 = ((uint64_t)(w.z >> 6) << 26) | (uint64_t)(w.w >> 6);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincos2pi_from_bits(k2, s, c);
    z0 = rad * c;
    z1 = rad * s;
}

// same pair with the radius logarithm from the shared-memory copy of g_logtab
__device__ __forceinline__ void draw_normal_pair_tab(uint32_t k0, uint32_t k1, uint64_t gid,
                                                     int block, uint32_t epoch, uint32_t p,
                                                     const double *tab, double &z0, double &z1) {
    u32x4 w = philox4x32_10(k0, k1, p, epoch, (uint32_t)gid,
                            CB2_TAG_BASIS | ((uint32_t)block << 8));
    const double u1 = u52(w.x, w.y);
    const uint64_t k2 = ((uint64_t)(w.z >> 6) << 26) | (uint64_t)(w.w >> 6);
    const double rad = sqrt(-2.0 * log_tab(u1, tab));
    double s, c;
    sincos2pi_from_bits(k2, s, c);
    z0 = rad * c;
    z1 = rad * s;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
