// =============================================================================
// dem_kernels.cuh -- hand-written sm_100a kernels of the SMC granular DEM step.
//
// One step = one replay of the same CUDA graph (see dem_engine.cu for the launch order):
//   k_step_begin       1 thread    travel bookkeeping -> "rebuild this step?", Multicore grid of the step
//                                  (ChBroadphase.cpp:143-208), search grid, ping-pong buffer indices
//   -- only when the Verlet skin is used up (every kernel below returns at once otherwise) --
//   k_bin_count        N threads   sphere -> search cell, histogram with warp-aggregated atomics
//   k_scan_*           ncell       exclusive scan of the histogram (CSR cell starts), self-cleaning
//   k_scatter_perm     N threads   counting-sort permutation
//   k_gather_sorted    N threads   re-order pos / vel / history columns into cell order
//   k_build_list       N threads   27-cell scan -> candidate list (r_i + r_j + skin), sorted by stable id
//   -- every step --
//   k_force_integrate  N threads   exact narrowphase on the candidates (sphere_sphere, ChNarrowphasePRIMS.cpp:40-72;
//                                  box_sphere :269-313), Hertz/Hooke/... force law with the pair-keyed tangential
//                                  history (ChIterativeSolverMulticoreSMC.cpp:56-546), rolling/spinning resistance,
//                                  gravity and the time integrator -- fused, one pass over the state
// Everything that decides contact-pair membership or Multicore bin ids uses explicitly rounded fp64 intrinsics
// (__dmul_rn/__dadd_rn/__dsub_rn) so no FMA contraction can change a bit w.r.t. the Multicore arithmetic.  The
// force arithmetic itself is free to contract/reassociate: its bar is 1e-9 relative, not bit identity.
// =============================================================================
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <type_traits>
#include "dem_types.h"

namespace demb200 {

// --------------------------------------------------------------------------------------------
// small fp64 vector helpers
// --------------------------------------------------------------------------------------------
struct V3 {
    double x, y, z;
};
__device__ __forceinline__ V3 mk(double x, double y, double z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator/(V3 a, double s) { return V3{a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double len(V3 a) { return sqrt(dot(a, a)); }
// Dot product with the reference's rounding sequence (simd_non.h:56-58): (x*x + y*y) + z*z, no contraction.
__device__ __forceinline__ double dot_rn(V3 a, V3 b) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)), __dmul_rn(a.z, b.z));
}
// Rotate / RotateT (src/chrono/multicore_math/real4.cpp:158-165), explicitly rounded
__device__ __forceinline__ V3 cross_rn(V3 a, V3 b) {
    return V3{__dsub_rn(__dmul_rn(a.y, b.z), __dmul_rn(a.z, b.y)), __dsub_rn(__dmul_rn(a.z, b.x), __dmul_rn(a.x, b.z)),
              __dsub_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x))};
}
__device__ __forceinline__ V3 rotate_rn(V3 v, double qw, V3 qv) {
    V3 c = cross_rn(qv, v);
    V3 t = V3{__dmul_rn(2.0, c.x), __dmul_rn(2.0, c.y), __dmul_rn(2.0, c.z)};
    V3 c2 = cross_rn(qv, t);
    return V3{__dadd_rn(__dadd_rn(v.x, __dmul_rn(qw, t.x)), c2.x), __dadd_rn(__dadd_rn(v.y, __dmul_rn(qw, t.y)), c2.y),
              __dadd_rn(__dadd_rn(v.z, __dmul_rn(qw, t.z)), c2.z)};
}

__device__ __forceinline__ V3 add_rn(V3 a, V3 b) { return V3{__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y), __dadd_rn(a.z, b.z)}; }
__device__ __forceinline__ V3 sub_rn(V3 a, V3 b) { return V3{__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z)}; }
__device__ __forceinline__ V3 scale_rn(double s, V3 a) { return V3{__dmul_rn(s, a.x), __dmul_rn(s, a.y), __dmul_rn(s, a.z)}; }
__device__ __forceinline__ V3 div_rn(V3 a, double s) { return V3{__ddiv_rn(a.x, s), __ddiv_rn(a.y, s), __ddiv_rn(a.z, s)}; }

// order-preserving map double <-> uint64 for atomicMin / atomicMax
__device__ __forceinline__ unsigned long long enc_ord(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_ord(unsigned long long u) {
    u = (u >> 63) ? (u & 0x7FFFFFFFFFFFFFFFull) : ~u;
    return __longlong_as_double((long long)u);
}
__host__ __device__ inline unsigned long long enc_ord_h(double d) {
    unsigned long long u;
    memcpy(&u, &d, 8);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

__device__ __forceinline__ double sphere_mass(const Params& P, double r) {
    return __dmul_rn(P.mass_coef, __dmul_rn(__dmul_rn(r, r), r));
}

// ---- 64-byte velocity record access (4 x 16-byte transactions) ----
struct VelVal {
    V3 v, w;
    unsigned sid, meta;
    unsigned long long amask;
};
// 256-bit global accesses (LDG.E.ENL2.256 / STG.E.ENL2.256, sm_100): a gather of a 32-byte record is ONE request of
// 32 L1 wavefronts instead of two; the force kernel is bound by the L1 data pipe (ncu r01o: 74 % of its peak), not DRAM.
// ld256: arrays that the running kernel only reads (the compiler may schedule it freely); ld256v: locations the same
// thread also writes (history rows) -- ordered with respect to st256.
// Cache-policy qualifiers of the force kernel's access classes (PTX .level::eviction_priority; "" = default policy).
// History rows and candidate ids are pure streams (each byte is used once per step by one thread); partner records are
// gathered ~14 times per step by neighbouring threads.
#ifndef DEMB200_Q_HLD
#define DEMB200_Q_HLD ""   /* history rows, load */
#endif
#ifndef DEMB200_Q_HST
#define DEMB200_Q_HST ""   /* history rows, store */
#endif
#ifndef DEMB200_Q_GLD
#define DEMB200_Q_GLD ""   /* partner position / velocity gathers */
#endif
#ifndef DEMB200_Q_NL
#define DEMB200_Q_NL ""    /* candidate list entries */
#endif
#ifndef DEMB200_Q_SST
#define DEMB200_Q_SST ""   /* new state of the sphere, store */
#endif
__device__ __forceinline__ double4 ld256(const void* p) {
    double4 r;
    asm("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double4 ld256g(const void* p) {  // gathered partner record
    double4 r;
    asm("ld.global" DEMB200_Q_GLD ".v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double4 ld256v(const void* p) {
    double4 r;
    asm volatile("ld.global" DEMB200_Q_HLD ".v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st256(void* p, double4 v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ void st256h(void* p, double4 v) {  // history row
    asm volatile("st.global" DEMB200_Q_HST ".v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ void st256s(void* p, double4 v) {  // new state of a sphere
    asm volatile("st.global" DEMB200_Q_SST ".v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ unsigned ld_nl(const uint32_t* p) {  // candidate list entry
    unsigned r;
    asm("ld.global" DEMB200_Q_NL ".u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ VelVal load_vel(const VelRec* __restrict__ a, size_t i) {
    const double4 q0 = ld256(a + i), q1 = ld256(reinterpret_cast<const char*>(a + i) + 32);
    VelVal r;
    r.v = mk(q0.x, q0.y, q0.z);
    r.w = mk(q0.w, q1.x, q1.y);
    r.sid = (unsigned)__double2loint(q1.z);
    r.meta = (unsigned)__double2hiint(q1.z);
    r.amask = (unsigned long long)__double_as_longlong(q1.w);
    return r;
}
__device__ __forceinline__ VelVal load_vel_g(const VelRec* __restrict__ a, size_t i) {
    const double4 q0 = ld256g(a + i), q1 = ld256g(reinterpret_cast<const char*>(a + i) + 32);
    VelVal r;
    r.v = mk(q0.x, q0.y, q0.z);
    r.w = mk(q0.w, q1.x, q1.y);
    r.sid = (unsigned)__double2loint(q1.z);
    r.meta = (unsigned)__double2hiint(q1.z);
    r.amask = (unsigned long long)__double_as_longlong(q1.w);
    return r;
}
__device__ __forceinline__ void store_vel(VelRec* a, size_t i, V3 v, V3 w, unsigned sid, unsigned meta,
                                          unsigned long long amask) {
    st256(a + i, make_double4(v.x, v.y, v.z, w.x));
    st256(reinterpret_cast<char*>(a + i) + 32,
          make_double4(w.y, w.z, __hiloint2double((int)meta, (int)sid), __longlong_as_double((long long)amask)));
}
__device__ __forceinline__ void store_vel_s(VelRec* a, size_t i, V3 v, V3 w, unsigned sid, unsigned meta,
                                          unsigned long long amask) {
    st256s(a + i, make_double4(v.x, v.y, v.z, w.x));
    st256s(reinterpret_cast<char*>(a + i) + 32,
          make_double4(w.y, w.z, __hiloint2double((int)meta, (int)sid), __longlong_as_double((long long)amask)));
}
// history record: (disp xyz, packed key | steps << 32); the key is only meaningful in the staging buffers
__device__ __forceinline__ double pack_key(unsigned key, unsigned steps) { return __hiloint2double((int)steps, (int)key); }
__device__ __forceinline__ unsigned rec_key(double w) { return (unsigned)__double2loint(w); }
__device__ __forceinline__ unsigned rec_steps(double w) { return (unsigned)__double2hiint(w); }

#ifndef DEMB200_PF_L2
#define DEMB200_PF_L2 0
#endif
__device__ __forceinline__ void prefetch_l1(const void* p) {
#if DEMB200_PF_L2
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}

// Fast reciprocal / reciprocal square root for normal, positive, well-scaled arguments: hardware seed
// (MUFU.RCP64H / MUFU.RSQ64H, ~2^-23) + one cubically convergent correction, error <= ~1 ulp, no special-case branch.
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    e = fma(e, e, e);
    return fma(y, e, y);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = y * y;
    const double e = fma(x, -t, 1.0);
    const double p = fma(e, 0.375, 0.5);
    const double q = y * e;
    return fma(p, q, y);
}

// --------------------------------------------------------------------------------------------
// bounding box of all sphere AABBs (init / after set_state); per step it is fused into k_force_integrate
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_bbox_commit(double mnx, double mny, double mnz, double mxx, double mxy, double mxz,
                                                  unsigned long long* bbox) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mnz = fmin(mnz, __shfl_xor_sync(0xffffffffu, mnz, o));
        mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        mxz = fmax(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
    }
    if ((threadIdx.x & 31) == 0) {
        // Only touch the atomic when the warp actually improves the running extreme: with walls around the bed the
        // accumulator starts at the wall AABB and no atomic is ever issued.
        unsigned long long e;
        e = enc_ord(mnx); if (e < bbox[0]) atomicMin(&bbox[0], e);
        e = enc_ord(mny); if (e < bbox[1]) atomicMin(&bbox[1], e);
        e = enc_ord(mnz); if (e < bbox[2]) atomicMin(&bbox[2], e);
        e = enc_ord(mxx); if (e > bbox[3]) atomicMax(&bbox[3], e);
        e = enc_ord(mxy); if (e > bbox[4]) atomicMax(&bbox[4], e);
        e = enc_ord(mxz); if (e > bbox[5]) atomicMax(&bbox[5], e);
    }
}

__global__ void __launch_bounds__(256) k_bbox_reduce(Params P, Buffers B) {
    Ctrl& C = *B.ctrl;
    const double4* __restrict__ posr = B.pos[C.cur];
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    double mnx = CUDART_INF, mny = CUDART_INF, mnz = CUDART_INF, mxx = -CUDART_INF, mxy = -CUDART_INF, mxz = -CUDART_INF;
    if (i < P.N) {
        double4 p = posr[i];
        mnx = p.x - p.w; mny = p.y - p.w; mnz = p.z - p.w;
        mxx = p.x + p.w; mxy = p.y + p.w; mxz = p.z + p.w;
    }
    block_bbox_commit(mnx, mny, mnz, mxx, mxy, mxz, C.bbox);
    block_bbox_commit(mnx, mny, mnz, mxx, mxy, mxz, C.sbox);
}

// --------------------------------------------------------------------------------------------
// step control: runs as one thread at the head of every step
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void step_begin_body(const Params& P, const Buffers& B, Ctrl& C, const WallSet& WS);

// The control block is staged through shared memory by the whole warp: the serial part then works at shared-memory
// latency instead of paying a global round trip for every field it touches.
// `cond` != 0: handle of the conditional node of the step graph that holds the rebuild kernels; its body only runs in the
// steps that rebuild (no empty launches of N / 256-block grids in the other ~97 % of the steps).
__global__ void __launch_bounds__(32) k_step_begin(Params P, Buffers B, unsigned long long cond) {
    __shared__ Ctrl sC;
    __shared__ WallSet sW;  // read-only copy: the serial part below touches every wall, one global round trip each otherwise
    __shared__ unsigned s_err_in;
    static_assert(sizeof(Ctrl) % 8 == 0, "Ctrl is copied in 8-byte words");
    static_assert(offsetof(Ctrl, err) % 8 == 0 && offsetof(Ctrl, nsteps) == offsetof(Ctrl, err) + 8, "err owns its 8-byte word");
    if (blockIdx.x != 0)
        return;
    unsigned long long* g = reinterpret_cast<unsigned long long*>(B.ctrl);
    unsigned long long* l = reinterpret_cast<unsigned long long*>(&sC);
    constexpr unsigned kWords = sizeof(Ctrl) / 8;
    for (unsigned i = threadIdx.x; i < kWords; i += blockDim.x)
        l[i] = g[i];
    {
        static_assert(sizeof(WallSet) % 8 == 0, "WallSet is copied in 8-byte words");
        const unsigned long long* gw = reinterpret_cast<const unsigned long long*>(B.walls);
        unsigned long long* lw = reinterpret_cast<unsigned long long*>(&sW);
        for (unsigned i = threadIdx.x; i < sizeof(WallSet) / 8; i += blockDim.x)
            lw[i] = gw[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        s_err_in = sC.err;
        step_begin_body(P, B, sC, sW);
        if (cond)
            cudaGraphSetConditional((cudaGraphConditionalHandle)cond, sC.rebuild_now);
    }
    __syncthreads();
    // the word that holds `err` is not written back: error bits may be set concurrently by kernels of another stream (halo
    // overlap), and a plain store of the staged copy would lose them.  New bits of this kernel are OR-ed in instead.
    constexpr unsigned kErrWord = offsetof(Ctrl, err) / 8;
    for (unsigned i = threadIdx.x; i < kWords; i += blockDim.x)
        if (i != kErrWord)
            g[i] = l[i];
    if (threadIdx.x == 0 && sC.err != s_err_in)
        atomicOr(&B.ctrl->err, sC.err);
}

__device__ __forceinline__ void step_begin_body(const Params& P, const Buffers& B, Ctrl& C, const WallSet& WS) {
    if (C.sbox_pending) {  // the last rebuild measured the spheres' own box while binning them
        for (int k = 0; k < 6; k++)
            C.sbox[k] = C.sbox_next[k];
        C.sbox_pending = 0u;
    }
    // ---- Verlet bookkeeping: every sphere moved at most `travel` since the lists were built ----
    const double dx = sqrt(__longlong_as_double((long long)C.max_dx2));
    C.max_dx2 = 0ull;
    C.travel += dx;
    C.last_dx = dx;
    // positions the host replaced since the last step (k_import_state) moved the spheres as well: the largest such jump counts
    // like one more step's displacement.  A host that hands back what it was given (a co-simulation round trip) costs no rebuild.
    C.travel += sqrt(__longlong_as_double((long long)C.import_dx2));
    C.import_dx2 = 0ull;
    if (!P.skin_adaptive || !(C.skin > 0))
        C.skin = P.skin;
    const bool mesh_used_up = P.nT && !(C.travel + C.travel_mesh < 0.499 * P.skin_tri);
    const bool rebuild = (C.need_rebuild != 0) || (!P.external_rebuild && (!(C.travel < 0.499 * C.skin) || mesh_used_up));
    if (rebuild && P.skin_adaptive && C.nrebuilds >= 1) {
        if (C.max_cand + 6u > (unsigned)P.Kn)  // the lists are nearly full: never trade rebuilds for an overflow
            C.skin = fmax(C.skin / 1.3, P.skin);
        else if (C.since_rebuild < 12u && C.max_cand + 12u <= (unsigned)P.Kn)
            C.skin = fmin(1.3 * C.skin, P.skin_max);
        else if (C.since_rebuild > 60u)
            C.skin = fmax(C.skin / 1.3, P.skin);
    }
    if (rebuild) {
        C.max_cand = 0u;
        C.n_bnd = 0u;
    }
    C.since_rebuild = rebuild ? 0u : C.since_rebuild + 1u;
    if (C.nrebuilds >= 1)
        C.init_stage = 0u;  // the checkpoint history was consumed by the first rebuild
    if (P.external_rebuild && !rebuild && (!(C.travel < 0.5 * C.skin) || (P.nT && !(C.travel + C.travel_mesh < 0.5 * P.skin_tri))))
        atomicOr(&C.err, ERR_SKIN_EXCEEDED);  // the slab driver asked for the rebuild too late: lists may miss contacts

    // ---- bounding box of all shapes at the start of this step ----
    double mn[3], mx[3];
    for (int k = 0; k < 3; k++) {
        mn[k] = dec_ord(C.bbox[k]);
        mx[k] = dec_ord(C.bbox[3 + k]);
        if (WS.has_bb) {
            mn[k] = fmin(mn[k], WS.bb_min[k]);
            mx[k] = fmax(mx[k], WS.bb_max[k]);
        }
    }
    if (P.nT) {  // mesh triangles are shapes of the Multicore broadphase too (ChCollisionSystemMulticore.cpp:387-394)
        MeshSet& MS = *B.meshes;
        for (int m = 0; m < MS.n; m++) {
            for (int k = 0; k < 3; k++) {
                mn[k] = fmin(mn[k], dec_ord(MS.bb[m][k]));
                mx[k] = fmax(mx[k], dec_ord(MS.bb[m][3 + k]));
            }
            for (int k = 0; k < 6; k++)
                MS.wrench[m][k] = 0.0;
        }
    }
    // ---- Multicore grid: ChBroadphase::DetermineBoundingBox + ComputeTopLevelResolution (ChBroadphase.cpp:143-208)
    GridDev& G = C.mc;
    unsigned err = 0;
    for (int k = 0; k < 3; k++) {
        const double fraction = 1e-3;
        double size = __dsub_rn(mx[k], mn[k]);
        double lo = __dsub_rn(mn[k], __dmul_rn(fraction, size));
        double hi = __dadd_rn(mx[k], __dmul_rn(fraction, size));
        double diag = fabs(__dsub_rn(hi, lo));
        double bin = __ddiv_rn(diag, (double)P.bins[k]);
        G.origin[k] = lo;
        G.bin[k] = bin;
        G.inv[k] = __ddiv_rn(1.0, bin);
        if (!isfinite(lo) || !isfinite(hi))
            err |= ERR_NAN;
    }
    for (int w = 0; w < P.nW; w++)
        for (int k = 0; k < 3; k++) {
            G.wmin[w][k] = __dsub_rn(WS.w[w].amin[k], G.origin[k]);
            G.wmax[w][k] = __dsub_rn(WS.w[w].amax[k], G.origin[k]);
        }
    if (err)
        atomicOr(&C.err, err);

    // ---- search grid for the rebuild: cells no smaller than the candidate cut-off 2 rmax + skin, over the box of the spheres
    //      (as measured at the last rebuild, grown by what they can have moved since) and of the meshes -- not of the walls:
    //      positions outside the grid are clamped into its boundary cells, which is still exact, only crowded
    if (rebuild) {
        for (int k = 0; k < 3; k++) {
            mn[k] = dec_ord(C.sbox[k]) - C.travel;
            mx[k] = dec_ord(C.sbox[3 + k]) + C.travel;
        }
        if (P.nT) {
            const MeshSet& MS = *B.meshes;
            for (int m = 0; m < MS.n; m++)
                for (int k = 0; k < 3; k++) {
                    mn[k] = fmin(mn[k], dec_ord(MS.bb[m][k]));
                    mx[k] = fmax(mx[k], dec_ord(MS.bb[m][3 + k]));
                }
        }
        for (int k = 0; k < 3; k++)
            if (!(mx[k] >= mn[k]) || !isfinite(mn[k]) || !isfinite(mx[k]))
                mn[k] = mx[k] = 0.0;  // no spheres (cannot happen after initialize): a one-cell grid
        double e = (2.0 * P.rmax + C.skin) * (1.0 + 1e-9);
        double ext[3];
        for (int k = 0; k < 3; k++) {
            ext[k] = (mx[k] - mn[k]) + 2e-6 * e;
            if (!(ext[k] >= e))
                ext[k] = e;
        }
        int dim[3], tdim[3];
        for (int it = 0; it < 400; it++) {
            double tot = 1.0;
            for (int k = 0; k < 3; k++) {
                double d = floor(ext[k] / e);
                d = (d < 1.0) ? 1.0 : d;
                d = (d > 2097152.0) ? 2097152.0 : d;
                dim[k] = (int)d;
                tdim[k] = (dim[k] + kTile - 1) / kTile;
                tot *= P.tiled ? (double)(tdim[k] * kTile) : d;  // tiled: the cells of the partly filled border tiles count too
            }
            if (tot <= (double)P.cell_cap)
                break;
            e *= 1.05;
        }
        unsigned long long tot = P.tiled ? (unsigned long long)tdim[0] * tdim[1] * tdim[2] * (unsigned long long)kTileCells
                                         : (unsigned long long)dim[0] * dim[1] * dim[2];
        if (tot > P.cell_cap) {  // not reachable (400 x 5 % growth), but never index out of bounds
            dim[0] = dim[1] = dim[2] = 1;
            tdim[0] = tdim[1] = tdim[2] = 1;
            tot = P.tiled ? (unsigned long long)kTileCells : 1ull;
        }
        for (int k = 0; k < 3; k++) {
            C.s_org[k] = mn[k] - 1e-6 * e;
            C.s_inv[k] = (double)dim[k] / ext[k];
            C.s_dim[k] = dim[k];
            C.t_dim[k] = tdim[k];
        }
        C.s_ncell = (unsigned)tot;
        // slab mode: x (the slab axis) varies slowest, y fastest
        C.s_perm[0] = P.external_rebuild ? 1 : 0;
        C.s_perm[1] = P.external_rebuild ? 2 : 1;
        C.s_perm[2] = P.external_rebuild ? 0 : 2;
        for (int k = 0; k < 3; k++) {
            C.sbox_next[k] = enc_ord(CUDART_INF);
            C.sbox_next[3 + k] = enc_ord(-CUDART_INF);
        }
        C.sbox_pending = 1u;
        C.travel = 0.0;
        C.travel_mesh = 0.0;
        C.nrebuilds++;
    }
    C.rebuild_now = rebuild ? 1u : 0u;
    C.need_rebuild = 0u;
    C.rb_src = C.cur;
    C.f_src = C.cur ^ (rebuild ? 1u : 0u);
    C.cur = C.f_src ^ 1u;
    C.nsteps++;
    C.n_contacts = 0ull;
    C.pair_count = 0ull;
    // restart the running sphere bounding box for the positions this step will produce
    for (int k = 0; k < 3; k++) {
        C.bbox[k] = WS.has_bb ? enc_ord(WS.bb_min[k]) : enc_ord(CUDART_INF);
        C.bbox[3 + k] = WS.has_bb ? enc_ord(WS.bb_max[k]) : enc_ord(-CUDART_INF);
    }
    for (int w = 0; w < P.nW; w++)
        C.wall_force[w][0] = C.wall_force[w][1] = C.wall_force[w][2] = 0.0;
}

// A wall was moved by the host (dem_b200_set_wall_state): its displacement uses up Verlet skin like a sphere's.
__global__ void k_wall_moved(Buffers B, double dist) {
    if (threadIdx.x == 0 && blockIdx.x == 0)
        B.ctrl->travel += dist;
}
__global__ void k_mesh_moved(Buffers B, double dist) {
    if (threadIdx.x == 0 && blockIdx.x == 0)
        B.ctrl->travel_mesh += dist;
}

// HashMin of the sphere AABB lower corner, HashMax of the upper corner (ChCollisionUtils.h:44-60), computed on the
// origin-offset AABB exactly as OffsetAABB + f_Count_AABB_BIN_Intersection do.
__device__ __forceinline__ void sphere_aabb_offset(const double4& p, const double* org, double* amin, double* amax) {
    const double c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        amin[k] = __dsub_rn(__dsub_rn(c[k], p.w), org[k]);
        amax[k] = __dsub_rn(__dadd_rn(c[k], p.w), org[k]);
    }
}

// Parity output: Multicore bin range of every sphere at the start of the step (recording mode only).
__global__ void __launch_bounds__(256) k_record_bins(Params P, Buffers B) {
    const Ctrl& C = *B.ctrl;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N)
        return;
    const double4 p = B.pos[C.f_src][i];
    const unsigned sid = B.vel[C.f_src][i].sid;
    double amin[3], amax[3];
    sphere_aabb_offset(p, C.mc.origin, amin, amax);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        B.gmin[3 * (size_t)sid + k] = (int)floor(__dmul_rn(amin[k], C.mc.inv[k]));
        B.gmax[3 * (size_t)sid + k] = (int)ceil(__dmul_rn(amax[k], C.mc.inv[k])) - 1;
    }
}

// --------------------------------------------------------------------------------------------
// rebuild, part 1: search cell per sphere + histogram with warp-aggregated atomics
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ int cell_coord(double x, double org, double inv, int dim) {
    int c = (int)floor((x - org) * inv);
    return min(max(c, 0), dim - 1);
}
// Number of a search cell.  Plain grid: x fastest.  Tiled grid (Params::tiled): tiles of kTile^3 cells, tile by tile (x fastest
// over the tiles), x fastest inside a tile -- the spheres of one tile are then one contiguous run of the storage order.
__device__ __forceinline__ unsigned cell_index(const Params& P, const Ctrl& C, int cx, int cy, int cz) {
    if (!P.tiled) {
        const int c[3] = {cx, cy, cz};
        const int p0 = C.s_perm[0], p1 = C.s_perm[1], p2 = C.s_perm[2];
        return (unsigned)((c[p2] * C.s_dim[p1] + c[p1]) * C.s_dim[p0] + c[p0]);
    }
    const int tx = cx / kTile, ty = cy / kTile, tz = cz / kTile;
    const unsigned tile = (unsigned)((tz * C.t_dim[1] + ty) * C.t_dim[0] + tx);
    return tile * (unsigned)kTileCells + (unsigned)((((cz % kTile) * kTile) + (cy % kTile)) * kTile + (cx % kTile));
}

__global__ void __launch_bounds__(256) k_bin_count(Params P, Buffers B) {
    Ctrl& C = *B.ctrl;
    if (!C.rebuild_now)
        return;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < P.N;
    unsigned h = 0xFFFFFFFFu;
    {   // the spheres' own bounding box, for the search grid of the NEXT rebuild (warp reduction, atomics only when it grows)
        double4 p = make_double4(0, 0, 0, 0);
        if (valid)
            p = B.pos[C.rb_src][i];
        const double inf = CUDART_INF;
        block_bbox_commit(valid ? p.x - p.w : inf, valid ? p.y - p.w : inf, valid ? p.z - p.w : inf, valid ? p.x + p.w : -inf,
                          valid ? p.y + p.w : -inf, valid ? p.z + p.w : -inf, C.sbox_next);
    }
    if (valid) {
        const double4 p = B.pos[C.rb_src][i];
        const int cx = cell_coord(p.x, C.s_org[0], C.s_inv[0], C.s_dim[0]);
        const int cy = cell_coord(p.y, C.s_org[1], C.s_inv[1], C.s_dim[1]);
        const int cz = cell_coord(p.z, C.s_org[2], C.s_inv[2], C.s_dim[2]);
        h = cell_index(P, C, cx, cy, cz);
        B.cell[i] = h;
    }
    // warp-aggregated histogram update: lanes that fall in the same cell elect a leader that issues one atomic;
    // the others derive their rank from their position inside the group.
    unsigned active = __ballot_sync(0xffffffffu, valid);
    if (valid) {
        unsigned peers = __match_any_sync(active, h);
        int leader = __ffs(peers) - 1;
        unsigned lane = threadIdx.x & 31;
        unsigned base = 0;
        if ((int)lane == leader)
            base = atomicAdd(&B.cell_count[h], (unsigned)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        B.rank[i] = base + __popc(peers & ((1u << lane) - 1));
    }
}

// --------------------------------------------------------------------------------------------
// exclusive scan over the cell histogram: reduce-then-scan, 3 launches, tiles of 256 threads x 8 items
// --------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* total) {
    __shared__ unsigned warp_sums[kScanThreads / 32];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= (unsigned)o)
            inc += t;
    }
    if (lane == 31)
        warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned w = (lane < kScanThreads / 32) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < kScanThreads / 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= (unsigned)o)
                w += t;
        }
        if (lane < kScanThreads / 32)
            warp_sums[lane] = w;
    }
    __syncthreads();
    unsigned prefix = (wid > 0) ? warp_sums[wid - 1] : 0;
    if (total)
        *total = warp_sums[kScanThreads / 32 - 1];
    __syncthreads();
    return prefix + inc - v;
}

// which = 0: sphere histogram (cell_count -> cell_start); which = 1: (cell, triangle) pair histogram of the meshes
__global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(Buffers B, int which) {
    const Ctrl& C = *B.ctrl;
    if (!C.rebuild_now)
        return;
    const unsigned n = C.s_ncell;
    if (blockIdx.x * kScanTile >= n)
        return;
    const uint32_t* __restrict__ in = which ? B.tcell_count : B.cell_count;
    const unsigned base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    unsigned s = 0;
    if (base + kScanItems <= n) {
        const uint4* p = reinterpret_cast<const uint4*>(in + base);
        uint4 a = p[0], b = p[1];
        s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    } else {
        for (int k = 0; k < kScanItems; k++)
            if (base + k < n)
                s += in[base + k];
    }
    unsigned total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0)
        (which ? B.tblock_sums : B.block_sums)[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_sums(Buffers B, int which) {
    const Ctrl& C = *B.ctrl;
    if (!C.rebuild_now)
        return;
    const unsigned ntiles = (C.s_ncell + kScanTile - 1) / kScanTile;
    uint32_t* tile_sums = which ? B.tblock_sums : B.block_sums;
    __shared__ unsigned carry_s;
    if (threadIdx.x == 0)
        carry_s = 0;
    __syncthreads();
    for (unsigned base = 0; base < ntiles; base += kScanThreads) {
        unsigned i = base + threadIdx.x;
        unsigned v = (i < ntiles) ? tile_sums[i] : 0;
        unsigned total;
        unsigned ex = block_exclusive_scan(v, &total);
        unsigned carry = carry_s;
        if (i < ntiles)
            tile_sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0)
            carry_s = carry + total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(Params P, Buffers B, int which) {
    Ctrl& C = *B.ctrl;
    if (!C.rebuild_now)
        return;
    const unsigned n = C.s_ncell;
    if (blockIdx.x * kScanTile >= n)
        return;
    uint32_t* in = which ? B.tcell_count : B.cell_count;
    uint32_t* __restrict__ out = which ? B.tcell_start : B.cell_start;
    const unsigned base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    unsigned v[kScanItems];
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    unsigned ex = block_exclusive_scan(s, nullptr) + (which ? B.tblock_sums : B.block_sums)[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) {
            out[base + k] = ex;
            // sphere histogram: self-cleaning (ready for the next rebuild, no memset node in the graph); the triangle
            // histogram is counted back down to zero by k_tri_fill
            if (!which)
                in[base + k] = 0;
        }
        ex += v[k];
    }
    if (!which) {
        if (blockIdx.x == 0 && threadIdx.x == 0)
            out[n] = P.N;
    } else if (base < n && n <= base + kScanItems) {  // the thread that owns the last cell knows the total
        out[n] = ex;
        if (ex > P.tri_cap)
            atomicOr(&C.err, ERR_MESH_CAPACITY);
    }
}

// --------------------------------------------------------------------------------------------
// triangle meshes.  World-frame soup = body frame + rigid motion (TransformLocalToParent, utility.h:46, rounding
// pinned); at a rebuild every triangle is registered in the search cells its AABB, inflated by the largest reach of a
// sphere (r_max + skin/2), overlaps, so that a sphere only has to look at the triangle list of its own cell.
// --------------------------------------------------------------------------------------------
__global__ void k_mesh_begin(Buffers B, int m) {
    if (threadIdx.x || blockIdx.x)
        return;
    MeshSet& MS = *B.meshes;
    for (int k = 0; k < 3; k++) {
        MS.bb[m][k] = enc_ord(CUDART_INF);
        MS.bb[m][3 + k] = enc_ord(-CUDART_INF);
    }
}

__global__ void __launch_bounds__(256) k_mesh_transform(Buffers B, int m) {
    MeshSet& MS = *B.meshes;
    const MeshBody& M = MS.m[m];
    const unsigned t = M.tri_begin + blockIdx.x * blockDim.x + threadIdx.x;
    double mn[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, mx[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
    if (t < M.tri_end) {
        const V3 qv = mk(M.rot[1], M.rot[2], M.rot[3]);
        const V3 bp = mk(M.pos[0], M.pos[1], M.pos[2]);
        const double* l = B.tri_loc + 9 * (size_t)t;
        double* w = B.tri_w + 9 * (size_t)t;
#pragma unroll
        for (int v = 0; v < 3; v++) {
            const V3 g = add_rn(bp, rotate_rn(mk(l[3 * v], l[3 * v + 1], l[3 * v + 2]), M.rot[0], qv));
            w[3 * v] = g.x; w[3 * v + 1] = g.y; w[3 * v + 2] = g.z;
            mn[0] = fmin(mn[0], g.x); mn[1] = fmin(mn[1], g.y); mn[2] = fmin(mn[2], g.z);
            mx[0] = fmax(mx[0], g.x); mx[1] = fmax(mx[1], g.y); mx[2] = fmax(mx[2], g.z);
        }
    }
    block_bbox_commit(mn[0], mn[1], mn[2], mx[0], mx[1], mx[2], MS.bb[m]);
}

// cells reached by triangle t: AABB inflated by `reach`, culled by the distance of the cell centre to the triangle's plane
template <class Visit>
__device__ __forceinline__ void tri_cells(const Params& P, const Ctrl& C, const double* w, double reach, Visit visit) {
    const V3 A = mk(w[0], w[1], w[2]), Bv = mk(w[3], w[4], w[5]), Cv = mk(w[6], w[7], w[8]);
    int lo[3], hi[3];
    const double mn[3] = {fmin(A.x, fmin(Bv.x, Cv.x)) - reach, fmin(A.y, fmin(Bv.y, Cv.y)) - reach, fmin(A.z, fmin(Bv.z, Cv.z)) - reach};
    const double mx[3] = {fmax(A.x, fmax(Bv.x, Cv.x)) + reach, fmax(A.y, fmax(Bv.y, Cv.y)) + reach, fmax(A.z, fmax(Bv.z, Cv.z)) + reach};
    for (int k = 0; k < 3; k++) {
        lo[k] = cell_coord(mn[k], C.s_org[k], C.s_inv[k], C.s_dim[k]);
        hi[k] = cell_coord(mx[k], C.s_org[k], C.s_inv[k], C.s_dim[k]);
    }
    V3 n = cross(Bv - A, Cv - A);
    const double nl = len(n);
    const bool flat = nl > 0;
    if (flat)
        n = n / nl;
    const double cs[3] = {1.0 / C.s_inv[0], 1.0 / C.s_inv[1], 1.0 / C.s_inv[2]};
    // a point of the cell lies within half the cell diagonal of its centre; boundary cells are unbounded (clamping)
    const double slack = reach + 0.5 * sqrt(cs[0] * cs[0] + cs[1] * cs[1] + cs[2] * cs[2]) * (1.0 + 1e-9);
    for (int z = lo[2]; z <= hi[2]; z++)
        for (int y = lo[1]; y <= hi[1]; y++)
            for (int x = lo[0]; x <= hi[0]; x++) {
                const bool edge = x == 0 || y == 0 || z == 0 || x == C.s_dim[0] - 1 || y == C.s_dim[1] - 1 || z == C.s_dim[2] - 1;
                if (flat && !edge) {
                    const V3 c = mk(C.s_org[0] + (x + 0.5) * cs[0], C.s_org[1] + (y + 0.5) * cs[1], C.s_org[2] + (z + 0.5) * cs[2]);
                    if (fabs(dot(c - A, n)) > slack)
                        continue;
                }
                visit(cell_index(P, C, x, y, z));
            }
}

__global__ void __launch_bounds__(256) k_tri_count(Params P, Buffers B) {
    const Ctrl& C = *B.ctrl;
    if (!C.rebuild_now)
        return;
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.nT)
        return;
    const double reach = (P.rmax + 0.5 * P.skin_tri) * (1.0 + 1e-9);
    tri_cells(P, C, B.tri_w + 9 * (size_t)t, reach, [&](unsigned cell) { atomicAdd(&B.tcell_count[cell], 1u); });
}

__global__ void __launch_bounds__(256) k_tri_fill(Params P, Buffers B) {
    const Ctrl& C = *B.ctrl;
    if (!C.rebuild_now)
        return;
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.nT)
        return;
    const double reach = (P.rmax + 0.5 * P.skin_tri) * (1.0 + 1e-9);
    tri_cells(P, C, B.tri_w + 9 * (size_t)t, reach, [&](unsigned cell) {
        // counts back down to zero: the histogram is clean again for the next rebuild
        const unsigned at = B.tcell_start[cell] + atomicSub(&B.tcell_count[cell], 1u) - 1u;
        if (at < P.tri_cap)
            B.tcell_tri[at] = t;
    });
}

// --------------------------------------------------------------------------------------------
// rebuild, part 2: counting-sort permutation, then gather of the sphere records and history columns
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter_perm(Params P, Buffers B) {
    const Ctrl& C = *B.ctrl;
    if (!C.rebuild_now)
        return;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N)
        return;
    B.perm[B.cell_start[B.cell[i]] + B.rank[i]] = i;
}

__global__ void __launch_bounds__(256) k_gather_sorted(Params P, Buffers B) {
    Ctrl& C = *B.ctrl;
    if (!C.rebuild_now)
        return;
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.N)
        return;
    const unsigned a = C.rb_src, b = a ^ 1u;
    const unsigned src = B.perm[s];
    B.pos[b][s] = B.pos[a][src];
    const double2* vs = reinterpret_cast<const double2*>(B.vel[a] + src);
    double2* vd = reinterpret_cast<double2*>(B.vel[b] + s);
    const double2 q0 = vs[0], q1 = vs[1], q2 = vs[2], q3 = vs[3];
    vd[0] = q0; vd[1] = q1; vd[2] = q2; vd[3] = q3;  // masks are rewritten by k_build_list
    if (B.acc[0]) {
        const double2* as = reinterpret_cast<const double2*>(B.acc[a] + 6 * (size_t)src);
        double2* ad = reinterpret_cast<double2*>(B.acc[b] + 6 * (size_t)s);
        ad[0] = as[0]; ad[1] = as[1]; ad[2] = as[2];
    }
    if (!B.hist)
        return;
    // Live history records leave the old candidate slots: compact them, keyed by partner shape id, into the staging
    // column of the sphere's NEW storage slot; k_build_list drops them into the slots of the new candidate list.
    unsigned cnt = 0;
    if (C.init_stage && B.stage_init) {
        cnt = B.stage_cnt_init[src];
        for (unsigned k = 0; k < cnt; k++) {
            B.stage[(size_t)k * P.Np + s] = B.stage_init[(size_t)k * P.Np + src];
            if (B.stage_rel)
                B.stage_rel[(size_t)k * P.Np + s] = B.stage_rel_init[(size_t)k * P.Np + src];
        }
    } else {
        const unsigned meta = (unsigned)__double2hiint(q3.x);
        unsigned wmask = (meta >> 8) & 0xFFFFu;
        unsigned long long amask = (unsigned long long)__double_as_longlong(q3.y);
        while (wmask) {
            const int w = __ffs(wmask) - 1;
            wmask &= wmask - 1;
            const size_t si = (size_t)(P.Kn + w) * P.Np + src;
            if (cnt < (unsigned)P.K) {
                double4 r = B.hist[si];
                r.w = pack_key((unsigned)w, (unsigned)r.w);
                B.stage[(size_t)cnt * P.Np + s] = r;
                if (B.stage_rel)
                    B.stage_rel[(size_t)cnt * P.Np + s] = B.hrel[si];
            }
            cnt++;
        }
        while (amask) {
            const int k = __ffsll((long long)amask) - 1;
            amask &= amask - 1;
            const size_t si = (size_t)k * P.Np + src;
            if (cnt < (unsigned)P.K) {
                const unsigned jo = B.nl[si] & ~kEntryFlags;  // old storage slot of the partner, or kTriFlag | triangle
                double4 r = B.hist[si];
                r.w = pack_key((jo & kTriFlag) ? (unsigned)P.nW + (jo & ~kTriFlag) : P.shape_base + B.vel[a][jo].sid, (unsigned)r.w);
                B.stage[(size_t)cnt * P.Np + s] = r;
                if (B.stage_rel)
                    B.stage_rel[(size_t)cnt * P.Np + s] = B.hrel[si];
            }
            cnt++;
        }
        if (cnt > (unsigned)P.K) {
            atomicOr(&C.err, ERR_HISTORY_OVERFLOW);
            cnt = P.K;
        }
    }
    B.stage_cnt[s] = cnt;
}

// --------------------------------------------------------------------------------------------
// rebuild, part 3: Verlet candidate list.  One thread per sphere (cell order), 3x3 rows of 3 contiguous cells.
// --------------------------------------------------------------------------------------------
#ifndef DEMB200_LIST_THREADS
#define DEMB200_LIST_THREADS 64  /* r02: 32 -> 437 us, 64 -> 343, 128 -> 412, 256 -> 418 per rebuild of 1 M spheres */
#endif
constexpr int kListThreads = DEMB200_LIST_THREADS;

// Closest point of triangle ABC to P (Ericson, Real-time collision detection, p.141); plain arithmetic: only used
// to select candidates, with slack.
__device__ __forceinline__ void closest_on_triangle(V3 A, V3 Bv, V3 Cv, V3 Pp, V3& res) {
    const V3 AB = Bv - A, AC = Cv - A, AP = Pp - A;
    const double d1 = dot(AB, AP), d2 = dot(AC, AP);
    if (d1 <= 0 && d2 <= 0) { res = A; return; }
    const V3 BP = Pp - Bv;
    const double d3 = dot(AB, BP), d4 = dot(AC, BP);
    if (d3 >= 0 && d4 <= d3) { res = Bv; return; }
    const double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) { res = A + (d1 / (d1 - d3)) * AB; return; }
    const V3 CP = Pp - Cv;
    const double d5 = dot(AB, CP), d6 = dot(AC, CP);
    if (d6 >= 0 && d5 <= d6) { res = Cv; return; }
    const double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) { res = A + (d2 / (d2 - d6)) * AC; return; }
    const double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) { res = Bv + ((d4 - d3) / ((d4 - d3) + (d5 - d6))) * (Cv - Bv); return; }
    const double denom = 1.0 / (va + vb + vc);
    res = A + (vb * denom) * AB + (vc * denom) * AC;
}

__global__ void __launch_bounds__(kListThreads) k_build_list(Params P, Buffers B) {
    Ctrl& C = *B.ctrl;
    if (!C.rebuild_now)
        return;
    const unsigned s = blockIdx.x * kListThreads + threadIdx.x;
    if (s >= P.N)
        return;
    const double4* __restrict__ pos = B.pos[C.f_src];
    const VelRec* vel = B.vel[C.f_src];
    if (vel[s].meta & FLAG_GHOST) {  // ghosts take part in other spheres' lists only
        B.ncnt[s] = 0u;
        return;
    }
    const double4 me = pos[s];
    // Two fixed bodies never make a contact in Chrono::Multicore (both inactive: ChCollisionUtilsBroadphase.cpp:131-132),
    // and the walls / meshes belong to fixed bodies too: a fixed sphere only keeps non-fixed sphere candidates.
    const bool me_fixed = (vel[s].meta & FLAG_FIXED) != 0;
    const int cx = cell_coord(me.x, C.s_org[0], C.s_inv[0], C.s_dim[0]);
    const int cy = cell_coord(me.y, C.s_org[1], C.s_inv[1], C.s_dim[1]);
    const int cz = cell_coord(me.z, C.s_org[2], C.s_inv[2], C.s_dim[2]);
    unsigned tj[kMaxNeighbors], ts[kMaxNeighbors];
    unsigned short tc[kMaxNeighbors];  // tiled grid: (box cell, rank in cell) of the candidate, relative to this sphere's tile
    const int tx0 = (cx / kTile) * kTile, ty0 = (cy / kTile) * kTile, tz0 = (cz / kTile) * kTile;
    // axes in the order of the cell numbering: a0 runs fastest (the three cells c0-1 .. c0+1 of a row are one run of the storage
    // order on the plain grid), a2 slowest
    const int a0 = P.tiled ? 0 : C.s_perm[0], a1 = P.tiled ? 1 : C.s_perm[1], a2 = P.tiled ? 2 : C.s_perm[2];
    const int cc[3] = {cx, cy, cz};
    const int lo0 = max(cc[a0] - 1, 0), hi0 = min(cc[a0] + 1, C.s_dim[a0] - 1);
    int cnt = 0;
    bool overflow = false, has_ghost = false;
    if (!P.tiled) {
        // plain grid.  The scan is a chain of dependent gathers if written naively (cell range -> position -> id of every
        // accepted candidate); here the nine row ranges are fetched together, the positions of a row four at a time, and the
        // ids / flags of the accepted candidates in a second sweep, four at a time as well.
        unsigned rb[9], re[9];
        int nrow = 0;
        for (int e2 = -1; e2 <= 1; e2++) {
            const int v2 = cc[a2] + e2;
            if (v2 < 0 || v2 >= C.s_dim[a2])
                continue;
            for (int e1 = -1; e1 <= 1; e1++) {
                const int v1 = cc[a1] + e1;
                if (v1 < 0 || v1 >= C.s_dim[a1])
                    continue;
                const unsigned row = (unsigned)((v2 * C.s_dim[a1] + v1) * C.s_dim[a0]);
                rb[nrow] = B.cell_start[row + lo0];   // the three cells of a row are one run of the storage order
                re[nrow] = B.cell_start[row + hi0 + 1];
                nrow++;
            }
        }
        for (int r = 0; r < nrow; r++) {
            for (unsigned j0 = rb[r]; j0 < re[r]; j0 += 4u) {
                double4 pj[4];
#pragma unroll
                for (unsigned u = 0; u < 4u; u++)
                    pj[u] = (j0 + u < re[r]) ? pos[j0 + u] : me;
#pragma unroll
                for (unsigned u = 0; u < 4u; u++) {
                    const unsigned j = j0 + u;
                    if (j >= re[r] || j == s)
                        continue;
                    const double dx = pj[u].x - me.x, dy2 = pj[u].y - me.y, dz2 = pj[u].z - me.z;
                    const double d2 = dx * dx + dy2 * dy2 + dz2 * dz2;
                    const double rs = me.w + pj[u].w + C.skin;
                    if (d2 > rs * rs * (1.0 + 1e-12))
                        continue;
                    if (cnt < P.Kn)
                        tj[cnt++] = j;
                    else
                        overflow = true;
                }
            }
        }
        // ids and flags of the accepted candidates; two fixed bodies make no contact: drop those pairs here
        int keep = 0;
        for (int k0 = 0; k0 < cnt; k0 += 4) {
            unsigned sidk[4], metak[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (k0 + u < cnt) {
                    sidk[u] = vel[tj[k0 + u]].sid;
                    metak[u] = vel[tj[k0 + u]].meta;
                }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (k0 + u < cnt) {
                    if (me_fixed && (metak[u] & FLAG_FIXED))
                        continue;
                    has_ghost |= (metak[u] & FLAG_GHOST) != 0;
                    tj[keep] = tj[k0 + u];
                    ts[keep] = sidk[u];
                    tc[keep] = 0;
                    keep++;
                }
        }
        cnt = keep;
    } else {
    for (int e2 = -1; e2 <= 1; e2++) {
        const int v2 = cc[a2] + e2;
        if (v2 < 0 || v2 >= C.s_dim[a2])
            continue;
        for (int e1 = -1; e1 <= 1; e1++) {
            const int v1 = cc[a1] + e1;
            if (v1 < 0 || v1 >= C.s_dim[a1])
                continue;
            for (int v0 = lo0; v0 <= hi0; v0++) {  // tiled grid (a0, a1, a2 = x, y, z): cell by cell
                const int x = v0, y = v1, z = v2;
                const unsigned ci = cell_index(P, C, x, y, z);
                const unsigned jb = B.cell_start[ci], je = B.cell_start[ci + 1];
                const unsigned box = (unsigned)(((z - tz0 + 1) * kBox + (y - ty0 + 1)) * kBox + (x - tx0 + 1));
                for (unsigned j = jb; j < je; j++) {
                    if (j == s)
                        continue;
                    const double4 pj = pos[j];
                    const double dx = pj.x - me.x, dy2 = pj.y - me.y, dz2 = pj.z - me.z;
                    const double d2 = dx * dx + dy2 * dy2 + dz2 * dz2;
                    const double rs = me.w + pj.w + C.skin;
                    if (d2 > rs * rs * (1.0 + 1e-12))
                        continue;
                    if (me_fixed && (vel[j].meta & FLAG_FIXED))
                        continue;
                    if (cnt < P.Kn) {
                        tj[cnt] = j;
                        ts[cnt] = vel[j].sid;
                        tc[cnt] = (unsigned short)((box << kCodeRankBits) | min(j - jb, kCodeNoRank));
                        has_ghost |= (vel[j].meta & FLAG_GHOST) != 0;
                        cnt++;
                    } else {
                        overflow = true;
                    }
                }
            }
        }
    }
    }
    if (overflow)
        atomicOr(&C.err, ERR_NEIGHBOR_OVERFLOW);
    // insertion sort by stable id: the per-step contact order (= summation order, history column order) becomes
    // independent of the storage order.  Plain grid: one array of packed (stable id, storage slot) keys instead of three.
    if (!P.tiled) {
        unsigned long long tk[kMaxNeighbors];
        for (int a = 0; a < cnt; a++)
            tk[a] = ((unsigned long long)ts[a] << 32) | tj[a];
        for (int a = 1; a < cnt; a++) {
            const unsigned long long ka = tk[a];
            int b = a - 1;
            while (b >= 0 && tk[b] > ka) {
                tk[b + 1] = tk[b];
                b--;
            }
            tk[b + 1] = ka;
        }
        for (int a = 0; a < cnt; a++) {
            tj[a] = (unsigned)tk[a];
            ts[a] = (unsigned)(tk[a] >> 32);
        }
    } else
    for (int a = 1; a < cnt; a++) {
        const unsigned kj = tj[a], ks = ts[a];
        const unsigned short kc = tc[a];
        int b = a - 1;
        while (b >= 0 && ts[b] > ks) {
            tj[b + 1] = tj[b];
            ts[b + 1] = ts[b];
            tc[b + 1] = tc[b];
            b--;
        }
        tj[b + 1] = kj;
        ts[b + 1] = ks;
        tc[b + 1] = kc;
    }
    // mesh triangles within reach (r + skin/2; the mesh's own motion uses up skin like a wall's), ascending triangle
    // index, behind the sphere candidates.  ts[] keeps the staging key (shape id - shape_base, wrapping for triangles).
    int tcnt = 0;
    if (P.nT && B.meshes->enabled && !me_fixed) {
        const unsigned cell = cell_index(P, C, cx, cy, cz);
        const unsigned tb = B.tcell_start[cell], te = min(B.tcell_start[cell + 1], P.tri_cap);
        const double reach = me.w + 0.5 * P.skin_tri + 1e-9 * (me.w + fabs(me.x) + fabs(me.y) + fabs(me.z));
        const V3 c = mk(me.x, me.y, me.z);
        for (unsigned q = tb; q < te; q++) {
            const unsigned t = B.tcell_tri[q];
            const double* w = B.tri_w + 9 * (size_t)t;
            V3 cp;
            closest_on_triangle(mk(w[0], w[1], w[2]), mk(w[3], w[4], w[5]), mk(w[6], w[7], w[8]), c, cp);
            const V3 d = c - cp;
            if (dot(d, d) > reach * reach)
                continue;
            if (cnt + tcnt >= P.Kn) {
                overflow = true;
                continue;
            }
            int b = cnt + tcnt - 1;
            while (b >= cnt && (tj[b] & ~kTriFlag) > t) {
                tj[b + 1] = tj[b];
                ts[b + 1] = ts[b];
                b--;
            }
            tj[b + 1] = kTriFlag | t;
            ts[b + 1] = (unsigned)P.nW + t - P.shape_base;
            tcnt++;
        }
        if (overflow)
            atomicOr(&C.err, ERR_NEIGHBOR_OVERFLOW);
    }
    {
        const unsigned my_sid = vel[s].sid;
        for (int k = 0; k < cnt + tcnt; k++)
            B.nl[(size_t)k * P.Np + s] = tj[k] | ((k < cnt && ts[k] > my_sid) ? kHiFlag : 0u);
        if (P.tiled)
            for (int k = 0; k < cnt; k++)
                B.nl16[(size_t)k * P.Np + s] = (unsigned short)(tc[k] | (ts[k] > my_sid ? kCodeHi : 0u));
    }
    if ((unsigned)(cnt + tcnt) > C.max_cand)
        atomicMax(&C.max_cand, (unsigned)(cnt + tcnt));
    // walls the sphere can touch before the next rebuild: it moves less than skin/2 until then (walls only move
    // through dem_b200_set_wall_velocity, which requests a rebuild)
    unsigned wc = 0;
    {
        const double reach = me.w + 0.5 * C.skin + 1e-9 * (me.w + fabs(me.x) + fabs(me.y) + fabs(me.z));
        for (int w = 0; w < (me_fixed ? 0 : P.nW); w++) {
            const Wall& W = B.walls->w[w];
            bool near;
            if (W.type == WALL_BOX) {
                near = me.x - reach <= W.amax[0] && W.amin[0] <= me.x + reach && me.y - reach <= W.amax[1] &&
                       W.amin[1] <= me.y + reach && me.z - reach <= W.amax[2] && W.amin[2] <= me.z + reach;
            } else if (W.type == WALL_PLANE) {
                near = (me.x - W.pos[0]) * W.hdims[0] + (me.y - W.pos[1]) * W.hdims[1] + (me.z - W.pos[2]) * W.hdims[2] < reach;
            } else if (W.type == WALL_SPHERE) {
                const double dc = sqrt((me.x - W.pos[0]) * (me.x - W.pos[0]) + (me.y - W.pos[1]) * (me.y - W.pos[1]) +
                                       (me.z - W.pos[2]) * (me.z - W.pos[2]));
                near = (W.hdims[1] > 0) ? (dc - reach < W.hdims[0]) : (dc + reach > W.hdims[0]);
            } else if (W.type == WALL_ZCONE) {
                const double rho = sqrt((me.x - W.pos[0]) * (me.x - W.pos[0]) + (me.y - W.pos[1]) * (me.y - W.pos[1]));
                // |height above the surface| bounds the distance to it from above
                near = fabs((me.z - W.pos[2]) - W.hdims[0] * rho) < reach * sqrt(1.0 + W.hdims[0] * W.hdims[0]) + reach &&
                       me.z < W.hdims[2] + reach && me.z > W.hdims[1] - reach;
            } else {
                const double dxy = sqrt((me.x - W.pos[0]) * (me.x - W.pos[0]) + (me.y - W.pos[1]) * (me.y - W.pos[1]));
                near = (W.hdims[1] > 0) ? (dxy + reach > W.hdims[0]) : (dxy - reach < W.hdims[0]);
            }
            if (near)
                wc |= 1u << w;
        }
    }
    // (a sphere with a ghost among its candidates lies within the ghost cut of a slab face: it is one of the ghost SENDERS, which
    // k_mgpu_remap flags as the second-pass set of the direct halo; the flag set here only matters until that kernel has run)
    B.ncnt[s] = (unsigned)cnt | (has_ghost ? kBndFlag : 0u) | (wc << 8) | ((unsigned)tcnt << 24);
    // staged history -> slots of the new list (a record whose partner is no longer a candidate is dropped: that
    // contact has broken)
    if (B.hist) {
        unsigned long long amask = 0ull;
        unsigned wmask = 0u;
        const unsigned sc = B.stage_cnt[s];
        // The staged records come in the order of the old list (spheres by stable id, then facets): the same order as the new
        // list, so the search for a record's slot resumes where the last one ended (and wraps, for records staged in any other
        // order: checkpoints, migrants).  Four records are in flight at a time.
        const int nk = cnt + tcnt;
        int kp = 0;
        for (unsigned c0 = 0; c0 < sc; c0 += 4u) {
            double4 rr[4];
#pragma unroll
            for (unsigned u = 0; u < 4u; u++)
                if (c0 + u < sc)
                    rr[u] = B.stage[(size_t)(c0 + u) * P.Np + s];
#pragma unroll
            for (unsigned u = 0; u < 4u; u++) {
                if (c0 + u >= sc)
                    continue;
                const unsigned c = c0 + u;
                double4 r = rr[u];
                const unsigned key = rec_key(r.w);
                r.w = (double)rec_steps(r.w);  // in the candidate slots the 4th component is the step count as a double
                size_t di;
                if (key < (unsigned)P.nW) {
                    di = (size_t)(P.Kn + key) * P.Np + s;
                    wmask |= 1u << key;
                } else {
                    const unsigned ps = key - P.shape_base;  // triangles: wraps, as stored in ts[]
                    int k = kp;
                    while (k < nk && ts[k] != ps)
                        k++;
                    if (k == nk) {
                        k = 0;
                        while (k < kp && ts[k] != ps)
                            k++;
                        if (k == kp)
                            continue;
                    }
                    kp = k + 1;
                    di = (size_t)k * P.Np + s;
                    amask |= 1ull << k;
                }
                B.hist[di] = r;
                if (B.hrel)
                    B.hrel[di] = B.stage_rel[(size_t)c * P.Np + s];
            }
        }
        VelRec* vr = B.vel[C.f_src] + s;
        vr->meta = (vr->meta & 0xFFu) | (wmask << 8);
        vr->amask = amask;
    }
}

// --------------------------------------------------------------------------------------------
// contact force law, generic: ChIterativeSolverMulticoreSMC.cpp:56-546, bodies without orientation (spheres: the
// body frame can be taken parallel to the world frame), canonical orientation body1 = lower shape id.
// Used for wall contacts and for every model combination that has no specialised fast path.
// --------------------------------------------------------------------------------------------
struct Body {
    V3 pos, v, w;
    double mass;
};
struct Geom {
    V3 n, pt1, pt2;
    double depth, erad;
};
struct Hist {
    V3 disp;
    double dur, relvel0;
    bool isnew;
};
// what SetRecordingContactInfo keeps of a contact, seen from body 2 of the evaluation (generic law) / from the evaluating
// sphere (fast law): normal and tangential part of the force on it, rolling + spinning resistance torque on it (tr1: on
// body 1), v_rot, characteristic collision time
struct CInfoOut {
    V3 fn, ft, tr, tr1, vrot;
    double tc;
};
__device__ __forceinline__ void store_cinfo(double* dst, V3 fn, V3 ft, V3 tr, V3 vrot, double tc) {
    dst[0] = fn.x; dst[1] = fn.y; dst[2] = fn.z; dst[3] = ft.x; dst[4] = ft.y; dst[5] = ft.z;
    dst[6] = tr.x; dst[7] = tr.y; dst[8] = tr.z; dst[9] = vrot.x; dst[10] = vrot.y; dst[11] = vrot.z; dst[12] = tc;
}

template <bool HIST, bool ROLL>
__device__ __noinline__ void contact_force(const Params& P, const Comp& cm, const Body& b1, const Body& b2,
                                           const Geom& g, Hist& h, V3& F, V3& T1, V3& T2, CInfoOut* info = nullptr) {
    const double kPI = 3.141592653589793238462643383279;
    const double eps = 2.220446049250313e-16;
    const V3 pt1_loc = g.pt1 - b1.pos;
    const V3 pt2_loc = g.pt2 - b2.pos;
    const V3 vel1 = b1.v + cross(b1.w, pt1_loc);
    const V3 vel2 = b2.v + cross(b2.w, pt2_loc);
    const V3 relvel = vel2 - vel1;
    const double relvel_n_mag = dot(relvel, g.n);
    const V3 relvel_n = relvel_n_mag * g.n;
    const V3 relvel_t = relvel - relvel_n;

    const double m_eff = b1.mass * b2.mass / (b1.mass + b2.mass);
    const double delta_n = -g.depth;
    double relvel_init = fabs(relvel_n_mag);
    double t_contact = 0;
    double char_vel = P.char_vel;
    V3 delta_t = mk(0, 0, 0);

    if (P.tang_mode == 1) {  // OneStep
        delta_t = relvel_t * P.dt;
    } else if (HIST) {  // MultiStep; history lives on body 2 (= max id) -> "else" branch of :233-243
        delta_t = relvel_t * P.dt;
        if (h.isnew) {
            h.disp = mk(0, 0, 0);
            h.relvel0 = relvel_init;
            h.dur = 0;
        } else {
            h.dur += P.dt;
        }
        h.disp = h.disp - delta_t;
        h.disp = h.disp - dot(h.disp, g.n) * g.n;
        delta_t = -h.disp;
        relvel_init = (h.relvel0 < char_vel) ? char_vel : h.relvel0;
        t_contact = h.dur;
    }

    double kn = 0, kt = 0, gn = 0, gt = 0, kn_simple = 0, gn_simple = 0;
    switch (P.force_model) {
        case 0:  // Hooke
            if (P.use_mat_props) {
                double tmp_k = (16.0 / 15) * sqrt(g.erad) * cm.E_eff;
                char_vel = (P.tang_mode == 2) ? relvel_init : char_vel;
                double v2 = char_vel * char_vel;
                double loge = (cm.cr < eps) ? log(eps) : log(cm.cr);
                loge = (cm.cr > 1 - eps) ? log(1 - eps) : loge;
                double q = kPI / loge;
                double tmp_g = 1 + q * q;
                kn = tmp_k * pow(m_eff * v2 / tmp_k, 0.2);
                kt = kn;
                gn = sqrt(4 * m_eff * kn / tmp_g);
                gt = gn;
            } else {
                kn = cm.kn; kt = cm.kt; gn = m_eff * cm.gn; gt = m_eff * cm.gt;
            }
            kn_simple = kn;
            gn_simple = gn;
            break;
        case 1:  // Hertz
            if (P.use_mat_props) {
                double sqrt_Rd = sqrt(g.erad * delta_n);
                double Sn = 2 * cm.E_eff * sqrt_Rd;
                double St = 8 * cm.G_eff * sqrt_Rd;
                kn = (2.0 / 3.0) * Sn;
                kt = St;
                gn = cm.hertz_damp * sqrt(Sn * m_eff);
                gt = cm.hertz_damp * sqrt(St * m_eff);
            } else {
                double tmp = g.erad * sqrt(delta_n);
                kn = tmp * cm.kn; kt = tmp * cm.kt; gn = tmp * m_eff * cm.gn; gt = tmp * m_eff * cm.gt;
            }
            if (ROLL) {
                kn_simple = kn / sqrt(delta_n);
                gn_simple = gn / pow(delta_n, 0.25);
            }
            break;
        case 3:  // Flores
            if (P.use_mat_props) {
                double sqrt_Rd = sqrt(g.erad * delta_n);
                double Sn = 2 * cm.E_eff * sqrt_Rd;
                double St = 8 * cm.G_eff * sqrt_Rd;
                double cr = (cm.cr < 0.01) ? 0.01 : cm.cr;
                cr = (cr > 1.0 - eps) ? 1.0 - eps : cr;
                double loge = log(cr);
                double beta = loge / sqrt(loge * loge + kPI * kPI);
                char_vel = (P.tang_mode == 2) ? relvel_init : char_vel;
                kn = (2.0 / 3.0) * Sn;
                kt = (2.0 / 3.0) * St;
                gn = 8.0 * (1.0 - cr) * kn * delta_n / (5.0 * cr * char_vel);
                gt = -2 * sqrt(5.0 / 6) * beta * sqrt(St * m_eff);
            } else {
                double tmp = g.erad * sqrt(delta_n);
                kn = tmp * cm.kn; kt = tmp * cm.kt; gn = tmp * m_eff * cm.gn * delta_n; gt = tmp * m_eff * cm.gt;
            }
            if (ROLL) {
                kn_simple = kn / sqrt(delta_n);
                gn_simple = gn / pow(delta_n, 1.5);
            }
            break;
        default:  // PlainCoulomb
            if (P.use_mat_props) {
                double Sn = 2 * cm.E_eff * sqrt(delta_n);
                kn = (2.0 / 3.0) * Sn;
                gn = cm.hertz_damp * sqrt(Sn * m_eff);
            } else {
                double tmp = sqrt(delta_n);
                kn = tmp * cm.kn; gn = tmp * cm.gn;
            }
            if (ROLL) {
                kn_simple = kn / sqrt(delta_n);
                gn_simple = gn / pow(delta_n, 0.25);
            }
            break;
    }

    const double forceN_mag = kn * delta_n - gn * relvel_n_mag;
    V3 force;
    if (P.force_model == 2) {
        double relvel_t_mag = len(relvel_t);
        double forceT_mag = cm.mu * tanh(5.0 * relvel_t_mag) * forceN_mag;
        force = forceN_mag * g.n;
        if (relvel_t_mag >= P.min_slip)
            force = force - (forceT_mag / relvel_t_mag) * relvel_t;
    } else {
        V3 forceT_damp = gt * relvel_t;
        V3 forceT = kt * delta_t + forceT_damp;
        double forceT_mag = len(forceT);
        double forceT_slide = cm.mu * fabs(forceN_mag);
        if (forceT_mag > forceT_slide) {
            if (len(delta_t) > eps) {
                double ratio = forceT_slide / forceT_mag;
                forceT = forceT * ratio;
                if (HIST) {
                    delta_t = (forceT - forceT_damp) / kt;
                    h.disp = -delta_t;
                }
            } else {
                forceT = mk(0, 0, 0);
            }
        }
        force = forceN_mag * g.n - forceT;
    }

    V3 tq1 = -cross(pt1_loc, force);
    V3 tq2 = cross(pt2_loc, force);
    if (info) {
        info->fn = forceN_mag * g.n;
        info->ft = force - info->fn;
        info->tr = mk(0, 0, 0); info->tr1 = mk(0, 0, 0); info->vrot = mk(0, 0, 0);
        info->tc = 0.0;
    }
    const V3 tq1_f = tq1, tq2_f = tq2;

    if (ROLL) {
        double muRoll = cm.mu_roll, muSpin = cm.mu_spin;
        double d_coeff = gn_simple / (2.0 * m_eff * sqrt(kn_simple / m_eff));
        if (d_coeff < 1.0) {
            double t_collision = kPI * sqrt(m_eff / (kn_simple * (1 - d_coeff * d_coeff)));
            if (info)
                info->tc = t_collision;
            if (t_contact <= t_collision) {
                muRoll = 0.0;
                muSpin = 0.0;
            }
        }
        V3 v_rot = cross(b2.w, pt2_loc) - cross(b1.w, pt1_loc);
        V3 rel_o = b2.w - b1.w;
        if (info)
            info->vrot = v_rot;
        double lv = len(v_rot);
        if (lv > P.min_roll && muRoll > eps) {
            tq1 = tq1 + muRoll * cross(forceN_mag * pt1_loc, v_rot) / lv;
            tq2 = tq2 - muRoll * cross(forceN_mag * pt2_loc, v_rot) / lv;
        }
        double lo = len(rel_o);
        if (lo > P.min_spin && muSpin > eps) {
            double r1 = len(pt1_loc), r2 = len(pt2_loc);
            double xc = (r1 * r1 - r2 * r2) / (2 * (r1 + r2 - delta_n)) + 0.5 * (r1 + r2 - delta_n);
            double rc = r1 * r1 - xc * xc;
            rc = (rc < eps) ? eps : sqrt(rc);
            V3 ms = muSpin * rc * (dot(rel_o, forceN_mag * g.n) * g.n) / lo;
            tq1 = tq1 + ms;
            tq2 = tq2 - ms;
        }
        if (info) {
            info->tr = tq2 - tq2_f;
            info->tr1 = tq1 - tq1_f;
        }
    }
    switch (P.adhesion_model) {
        case 0: force = force - cm.adh * g.n; break;
        case 1: force = force - cm.adh_dmt * sqrt(g.erad) * g.n; break;
        default: force = force - cm.adh_perko * g.erad * g.n; break;
    }
    if (info)
        info->fn = force - info->ft;  // adhesion acts along the normal
    F = force;
    T1 = tq1;
    T2 = tq2;
}

// --------------------------------------------------------------------------------------------
// Fast path of the same law: Hertz (material properties OR user coefficients), constant adhesion, sphere
// against sphere.  Written in the frame of the evaluating sphere ("a" = me, "b" = partner, n from a to b).  All
// expressions are odd/even under the exchange a <-> b with IEEE-exact sign symmetry (products are formed with
// commutative roundings), so both partners obtain bit-identical magnitudes and equal-and-opposite forces, and their
// two copies of the tangential history stay bit-identical.  The stored displacement is in canonical orientation
// (body 1 = lower shape id, ChIterativeSolverMulticoreSMC.cpp:233-243): disp_ab = sgn * disp_canonical.
// Divisions / square roots of the reference are replaced by rcp / rsqrt forms (<= 2 ulp); the bar is 1e-9.
// --------------------------------------------------------------------------------------------
template <bool HIST, bool ROLL, bool MATPROPS>
__device__ __forceinline__ void sphere_contact_fast(const Params& P, const Comp& cm, V3 n, double dist, double ra,
                                                    double rb, V3 va, V3 wa, V3 vb, V3 wb, double ma, double mb,
                                                    bool a_is_body1, V3& disp, double& steps, bool isnew, V3& F_me,
                                                    V3& T_me, CInfoOut* info = nullptr) {
    const double eps = 2.220446049250313e-16;
    const double radSum = __dadd_rn(ra, rb);
    const double delta_n = radSum - dist;
    // equal radii (every pair of a monodisperse packing; the test is symmetric in a <-> b, so both partners take the
    // same route): R* = r/2 and m* = m/2 without the two reciprocals
    double erad, m_eff;
    if (ra == rb) {
        erad = 0.5 * ra;
        m_eff = 0.5 * ma;
    } else {
        erad = __dmul_rn(ra, rb) * fast_rcp(radSum);
        m_eff = __dmul_rn(ma, mb) * fast_rcp(__dadd_rn(ma, mb));
    }
    // velocity of b's contact point minus a's:  (vb + wb x (-n rb)) - (va + wa x (n ra))
    const V3 wsum = mk(__dadd_rn(__dmul_rn(ra, wa.x), __dmul_rn(rb, wb.x)), __dadd_rn(__dmul_rn(ra, wa.y), __dmul_rn(rb, wb.y)),
                       __dadd_rn(__dmul_rn(ra, wa.z), __dmul_rn(rb, wb.z)));
    const V3 wxn = cross(wsum, n);
    const V3 relvel = (vb - va) - wxn;
    const double vn = dot(relvel, n);
    const V3 relvel_t = relvel - vn * n;

    V3 delta_t = mk(0, 0, 0);
    if (HIST) {
        delta_t = relvel_t * P.dt;
        if (isnew) {
            disp = mk(0, 0, 0);
            steps = 0.0;
        } else {
            steps += 1.0;
        }
        disp = disp - delta_t;
        disp = disp - dot(disp, n) * n;
        delta_t = -disp;
    } else if (P.tang_mode == 1) {
        delta_t = relvel_t * P.dt;
    }

    double kn, kt, gn, gt;
    if (MATPROPS) {  // ChIterativeSolverMulticoreSMC.cpp:277-290
        const double x = erad * delta_n;
        const double sqrt_Rd = x * fast_rsqrt(x);
        const double Sn = 2 * cm.E_eff * sqrt_Rd;
        const double St = 8 * cm.G_eff * sqrt_Rd;
        kn = (2.0 / 3.0) * Sn;
        kt = St;
        const double y = Sn * m_eff;
        gn = cm.hertz_damp * (y * fast_rsqrt(y));
        gt = gn * cm.gt_ratio;
    } else {  // user coefficients (:291-297) -- the model every Chrono::Dem setter (SetKn_SPH2SPH ...) maps to
        const double tmp = erad * (delta_n * fast_rsqrt(delta_n));
        kn = tmp * cm.kn;
        kt = tmp * cm.kt;
        gn = tmp * m_eff * cm.gn;
        gt = tmp * m_eff * cm.gt;
    }

    const double fN = kn * delta_n - gn * vn;
    const V3 fT_damp = gt * relvel_t;
    V3 fT = kt * delta_t + fT_damp;
    const double ft2 = dot(fT, fT);
    const double slide = cm.mu * fabs(fN);
    if (ft2 > slide * slide) {
        if (dot(delta_t, delta_t) > eps * eps) {
            const double ratio = slide * fast_rsqrt(ft2);
            fT = fT * ratio;
            if (HIST) {
                delta_t = (fT - fT_damp) * fast_rcp(kt);
                disp = -delta_t;
            }
        } else {
            fT = mk(0, 0, 0);
        }
    }
    // force on b = fN n - fT; on a (me) the opposite.  Torque on a: -(n ra) x (fN n - fT) = ra (n x fT)
    V3 Fb = fN * n - fT;
    V3 Ta = ra * cross(n, fT);
    const V3 Ta_f = Ta;
    if (info) {
        info->ft = fT;
        info->vrot = mk(0, 0, 0);
        info->tc = 0.0;
    }

    if (ROLL) {
        const double kPI = 3.141592653589793238462643383279;
        double muRoll = cm.mu_roll, muSpin = cm.mu_spin;
        // contact-duration gate (ChIterativeSolverMulticoreSMC.cpp:484-491) with reciprocal square roots instead of the five
        // square roots and four divisions of the reference expression:
        //   kn_simple = kn / sqrt(d), gn_simple = gn / d^(1/4), d_coeff = gn_simple / (2 m sqrt(kn_simple / m)),
        //   t_collision = pi sqrt(m / (kn_simple (1 - d_coeff^2)))
        const double ir_d = fast_rsqrt(delta_n);           // d^(-1/2)
        const double kn_simple = kn * ir_d;
        const double gn_simple = gn * fast_rsqrt(delta_n * ir_d);  // gn d^(-1/4)
        const double inv_m = fast_rcp(m_eff);
        const double d_coeff = 0.5 * gn_simple * inv_m * fast_rsqrt(kn_simple * inv_m);
        if (d_coeff < 1.0) {
            const double t_collision = kPI * fast_rsqrt(kn_simple * (1 - d_coeff * d_coeff) * inv_m);
            if (info)
                info->tc = t_collision;
            const double t_contact = HIST ? steps * P.dt : 0.0;
            if (t_contact <= t_collision) {
                muRoll = 0.0;
                muSpin = 0.0;
            }
        }
        // v_rot = wb x (-n rb) - wa x (n ra) = -(wsum x n)
        const V3 v_rot = -wxn;
        if (info)
            info->vrot = v_rot;
        const V3 rel_o = wb - wa;
        const double lv2 = dot(v_rot, v_rot);
        if (lv2 > P.min_roll * P.min_roll && muRoll > eps)
            Ta = Ta + (muRoll * fN * ra * fast_rsqrt(lv2)) * cross(n, v_rot);
        const double lo = len(rel_o);
        if (lo > P.min_spin && muSpin > eps) {
            // contact-circle radius, evaluated with body 1 = lower shape id as the reference does
            const double r1 = a_is_body1 ? ra : rb, r2 = a_is_body1 ? rb : ra;
            const double xc = (r1 * r1 - r2 * r2) / (2 * (r1 + r2 - delta_n)) + 0.5 * (r1 + r2 - delta_n);
            double rc = r1 * r1 - xc * xc;
            rc = (rc < eps) ? eps : sqrt(rc);
            Ta = Ta + (muSpin * rc * (dot(rel_o, n) * fN) / lo) * n;
        }
    }
    Fb = Fb - cm.adh * n;
    F_me = -Fb;
    T_me = Ta;
    if (info) {
        info->fn = (cm.adh - fN) * n;
        info->tr = Ta - Ta_f;
    }
}

// box_sphere: ChNarrowphasePRIMS.cpp:269-313 with snap_to_box (ChCollisionUtils.h:546-563); rounding pinned.
__device__ __forceinline__ bool box_sphere_dev(const Wall& W, V3 pos2, double r2, Geom& g) {
    const V3 qv = mk(W.rot[1], W.rot[2], W.rot[3]);
    const V3 nqv = mk(-W.rot[1], -W.rot[2], -W.rot[3]);
    const V3 bp = mk(W.pos[0], W.pos[1], W.pos[2]);
    V3 sp = rotate_rn(mk(__dsub_rn(pos2.x, bp.x), __dsub_rn(pos2.y, bp.y), __dsub_rn(pos2.z, bp.z)), W.rot[0], nqv);
    V3 bx = sp;
    unsigned code = 0;
    if (fabs(bx.x) > W.hdims[0]) { code |= 1; bx.x = (bx.x > 0) ? W.hdims[0] : -W.hdims[0]; }
    if (fabs(bx.y) > W.hdims[1]) { code |= 2; bx.y = (bx.y > 0) ? W.hdims[1] : -W.hdims[1]; }
    if (fabs(bx.z) > W.hdims[2]) { code |= 4; bx.z = (bx.z > 0) ? W.hdims[2] : -W.hdims[2]; }
    V3 delta = mk(__dsub_rn(sp.x, bx.x), __dsub_rn(sp.y, bx.y), __dsub_rn(sp.z, bx.z));
    double dist2 = dot_rn(delta, delta);
    if (dist2 >= __dmul_rn(r2, r2) || dist2 <= (double)1e-12f)
        return false;
    double dist = sqrt(dist2);
    g.depth = __dsub_rn(dist, r2);
    V3 dn = mk(__ddiv_rn(delta.x, dist), __ddiv_rn(delta.y, dist), __ddiv_rn(delta.z, dist));
    g.n = rotate_rn(dn, W.rot[0], qv);
    V3 p1 = rotate_rn(bx, W.rot[0], qv);
    g.pt1 = mk(__dadd_rn(bp.x, p1.x), __dadd_rn(bp.y, p1.y), __dadd_rn(bp.z, p1.z));
    g.pt2 = mk(__dsub_rn(pos2.x, __dmul_rn(g.n.x, r2)), __dsub_rn(pos2.y, __dmul_rn(g.n.y, r2)),
               __dsub_rn(pos2.z, __dmul_rn(g.n.z, r2)));
    g.erad = ((code != 1) && (code != 2) && (code != 4)) ? r2 * 0.1 / (r2 + 0.1) : r2;
    return true;
}

// Infinite plane (Chrono::Dem BC plane): contact iff signed distance < r; face contact, eff. radius r.
__device__ __forceinline__ bool plane_sphere_dev(const Wall& W, V3 pos2, double r2, Geom& g) {
    V3 n = mk(W.hdims[0], W.hdims[1], W.hdims[2]);
    V3 d = pos2 - mk(W.pos[0], W.pos[1], W.pos[2]);
    double dist = dot(d, n);
    if (dist >= r2)
        return false;
    g.n = n;
    g.depth = dist - r2;
    g.pt1 = pos2 - dist * n;
    g.pt2 = pos2 - r2 * n;
    g.erad = r2;
    return true;
}

// Z-axis cylinder (Chrono::Dem CreateBCCylinderZ, ChSystemDem.h:228; force form of addBCForces_Zcyl,
// ChDemBoundaryConditions.cuh): hdims = (radius, side): side +1 = spheres inside (normal towards the axis),
// side -1 = spheres outside.  Face contact, eff. radius r.
__device__ __forceinline__ bool zcyl_sphere_dev(const Wall& W, V3 pos2, double r2, Geom& g) {
    const double dx = pos2.x - W.pos[0], dy = pos2.y - W.pos[1];
    const double d = sqrt(dx * dx + dy * dy);
    const double Rc = W.hdims[0];
    const bool inside = W.hdims[1] > 0;
    const double gap = inside ? (Rc - d) : (d - Rc);  // distance from the sphere centre to the wall surface
    if (gap >= r2 || !(d > 1e-12 * Rc))
        return false;
    const double s = inside ? -1.0 : 1.0;
    g.n = mk(s * dx / d, s * dy / d, 0.0);  // from the wall towards the sphere
    g.depth = gap - r2;
    g.pt1 = pos2 - gap * g.n;
    g.pt2 = pos2 - r2 * g.n;
    g.erad = r2;
    return true;
}

// Ball (Chrono::Dem CreateBCSphere, ChSystemDem.h:206; force form addBCForces_Sphere_*, ChDemBoundaryConditions.cuh:155-215):
// hdims = (radius, side).  side +1: obstacle -- exactly Multicore's sphere_sphere against a fixed sphere (rounding pinned,
// ChNarrowphasePRIMS.cpp:40-72); side -1: cavity (spheres live inside the ball), concave effective radius.
__device__ __forceinline__ bool ball_sphere_dev(const Wall& W, V3 pos2, double r2, Geom& g) {
    const double Rb = W.hdims[0];
    const V3 c = mk(W.pos[0], W.pos[1], W.pos[2]);
    const V3 delta = sub_rn(pos2, c);
    const double d2 = dot_rn(delta, delta);
    if (W.hdims[1] > 0) {
        const double rs = __dadd_rn(Rb, r2);
        if (d2 >= __dmul_rn(rs, rs) || d2 < 1e-12)
            return false;
        const double dist = sqrt(d2);
        g.n = div_rn(delta, dist);
        g.pt1 = add_rn(c, scale_rn(Rb, g.n));
        g.pt2 = sub_rn(pos2, scale_rn(r2, g.n));
        g.depth = __dsub_rn(dist, rs);
        g.erad = __ddiv_rn(__dmul_rn(Rb, r2), rs);
        return true;
    }
    const double dist = sqrt(d2);
    const double gap = Rb - dist;  // distance from the sphere centre to the shell
    if (gap >= r2 || !(dist > 1e-12 * Rb) || !(Rb > r2))
        return false;
    g.n = delta * (-1.0 / dist);  // from the shell towards the sphere = towards the centre
    g.depth = gap - r2;
    g.pt1 = pos2 - gap * g.n;
    g.pt2 = pos2 - r2 * g.n;
    g.erad = Rb * r2 / (Rb - r2);
    return true;
}

// Cone about the z axis (Chrono::Dem CreateBCConeZ, ChSystemDem.h:212; geometry of addBCForces_ZCone_frictionless,
// ChDemBoundaryConditions.cuh:217-291): surface z - tip.z = slope * rho, active for hmin < z < hmax; closest point along
// the generator line below / above the sphere.  One-sided: rot[0] = +1 keeps spheres above the surface (hopper), -1 below.
__device__ __forceinline__ bool zcone_sphere_dev(const Wall& W, V3 pos2, double r2, Geom& g) {
    if (pos2.z >= W.hdims[2] || pos2.z <= W.hdims[1])
        return false;
    const V3 rel = pos2 - mk(W.pos[0], W.pos[1], W.pos[2]);
    const double rho = sqrt(rel.x * rel.x + rel.y * rel.y);
    const V3 l = mk(rel.x, rel.y, W.hdims[0] * rho);  // from the tip along the generator under the sphere
    const double ll = dot(l, l);
    if (!(ll > 0))
        return false;
    const V3 cv = rel - (dot(rel, l) / ll) * l;  // from the surface to the sphere centre
    const double dist = len(cv);
    const double side = rel.z - W.hdims[0] * rho;  // > 0: above the surface
    if (dist >= r2 || !(dist > 0) || side * W.rot[0] <= 0)
        return false;
    g.n = cv / dist;
    g.depth = dist - r2;
    g.pt1 = pos2 - cv;
    g.pt2 = pos2 - r2 * g.n;
    g.erad = r2;
    return true;
}

// snap_to_triangle (ChCollisionUtilsPRIMS.cpp:41-106) and triangle_sphere (ChNarrowphasePRIMS.cpp:379-437), rounding
// pinned: the hit / no-hit decision defines the contact-pair set and must not depend on FMA contraction.
__device__ __forceinline__ bool snap_to_triangle_rn(V3 A, V3 Bv, V3 Cv, V3 Pp, V3& res) {
    const V3 AB = sub_rn(Bv, A), AC = sub_rn(Cv, A), AP = sub_rn(Pp, A);
    const double d1 = dot_rn(AB, AP), d2 = dot_rn(AC, AP);
    if (d1 <= 0 && d2 <= 0) { res = A; return true; }
    const V3 BP = sub_rn(Pp, Bv);
    const double d3 = dot_rn(AB, BP), d4 = dot_rn(AC, BP);
    if (d3 >= 0 && d4 <= d3) { res = Bv; return true; }
    const double vc = __dsub_rn(__dmul_rn(d1, d4), __dmul_rn(d3, d2));
    if (vc <= 0 && d1 >= 0 && d3 <= 0) {
        res = add_rn(A, scale_rn(__ddiv_rn(d1, __dsub_rn(d1, d3)), AB));
        return true;
    }
    const V3 CP = sub_rn(Pp, Cv);
    const double d5 = dot_rn(AB, CP), d6 = dot_rn(AC, CP);
    if (d6 >= 0 && d5 <= d6) { res = Cv; return true; }
    const double vb = __dsub_rn(__dmul_rn(d5, d2), __dmul_rn(d1, d6));
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
        res = add_rn(A, scale_rn(__ddiv_rn(d2, __dsub_rn(d2, d6)), AC));
        return true;
    }
    const double va = __dsub_rn(__dmul_rn(d3, d6), __dmul_rn(d5, d4));
    const double e43 = __dsub_rn(d4, d3), e56 = __dsub_rn(d5, d6);
    if (va <= 0 && e43 >= 0 && e56 >= 0) {
        res = add_rn(Bv, scale_rn(__ddiv_rn(e43, __dadd_rn(e43, e56)), sub_rn(Cv, Bv)));
        return true;
    }
    const double denom = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(va, vb), vc));
    res = add_rn(add_rn(A, scale_rn(__dmul_rn(vb, denom), AB)), scale_rn(__dmul_rn(vc, denom), AC));
    return false;
}

__device__ __forceinline__ bool triangle_sphere_dev(V3 A, V3 Bv, V3 Cv, V3 pos2, double r2, Geom& g) {
    const V3 nx = cross_rn(sub_rn(Bv, A), sub_rn(Cv, A));  // triangle_normal, ChCollisionUtils.h:467-474
    const V3 nrm = div_rn(nx, sqrt(dot_rn(nx, nx)));
    const double h = dot_rn(sub_rn(pos2, A), nrm);
    if (h >= r2 || h <= 0)  // one-sided
        return false;
    V3 face;
    if (snap_to_triangle_rn(A, Bv, Cv, pos2, face)) {
        const V3 delta = sub_rn(pos2, face);
        const double dist2 = dot_rn(delta, delta);
        if (dist2 >= __dmul_rn(r2, r2) || dist2 <= (double)1e-12f)
            return false;
        const double dist = sqrt(dist2);
        g.n = div_rn(delta, dist);
        g.depth = __dsub_rn(dist, r2);
        g.erad = r2 * 0.1 / (r2 + 0.1);
    } else {
        g.n = nrm;
        g.depth = __dsub_rn(h, r2);
        g.erad = r2;
    }
    g.pt1 = face;
    g.pt2 = sub_rn(pos2, scale_rn(r2, g.n));
    return true;
}

// Sphere against the mesh triangles of its candidate list (slots first .. first+tc-1).  Body 1 = mesh body (lower id),
// body 2 = the sphere, exactly like a wall contact; every (sphere, triangle) pair is a contact of its own with its own
// history slot, as in Multicore (one shape per triangle).  The wrench on the mesh (force, torque about the body
// origin; reference: 6 atomics per contact, ChDemSMCtrimesh.cu:383-389) is summed over the lanes of the warp that hit
// the same mesh before it goes to memory.  Kept out of line: spheres near a mesh are a surface population.
struct MeshOut {
    V3 F, T;
    unsigned long long mask;
    unsigned n;
};

template <bool HIST, bool ROLL, bool REC>
__device__ __noinline__ void mesh_contacts(const Params& P, const Buffers& B, unsigned s, unsigned sid, double4 me, V3 v,
                                           V3 w, double my_mass, unsigned first, unsigned tc,
                                           unsigned long long amask_old, MeshOut& out) {
    MeshSet& MS = *B.meshes;
    Ctrl& C = *B.ctrl;
    out.F = mk(0, 0, 0);
    out.T = mk(0, 0, 0);
    out.mask = 0ull;
    out.n = 0;
    if (!MS.enabled)
        return;
    const V3 mpos = mk(me.x, me.y, me.z);
    double4* const hcol = HIST ? B.hist + s : nullptr;
    double* const rcol = (HIST && B.hrel) ? B.hrel + s : nullptr;
    for (unsigned k = 0; k < tc; k++) {
        const unsigned slot = first + k;
        const size_t hi = (size_t)slot * P.Np;
        const unsigned t = B.nl[hi + s] & ~kTriFlag;
        const double* tw = B.tri_w + 9 * (size_t)t;
        Geom g;
        if (!triangle_sphere_dev(mk(tw[0], tw[1], tw[2]), mk(tw[3], tw[4], tw[5]), mk(tw[6], tw[7], tw[8]), mpos, me.w, g))
            continue;
        out.n++;
        if (REC) {
            unsigned long long at = atomicAdd(&C.pair_count, 1ull);
            if (at < B.pair_cap)
                B.pairs[at] = ((unsigned long long)((unsigned)P.nW + t) << 32) | (unsigned long long)(P.shape_base + sid);
        }
        if (g.depth >= 0)
            continue;
        const int m = (int)B.tri_mesh[t];
        const MeshBody& M = MS.m[m];
        Hist h{mk(0, 0, 0), 0.0, 0.0, true};
        double steps = 0.0;
        if (HIST && ((amask_old >> slot) & 1ull)) {
            const double4 r = hcol[hi];
            h.disp = mk(r.x, r.y, r.z);
            steps = r.w;
            h.dur = steps * P.dt;
            if (rcol)
                h.relvel0 = rcol[hi];
            h.isnew = false;
            steps += 1.0;
        }
        Body b1{mk(M.pos[0], M.pos[1], M.pos[2]), mk(M.vel[0], M.vel[1], M.vel[2]), mk(M.omg[0], M.omg[1], M.omg[2]), M.mass};
        Body b2{mpos, v, w, my_mass};
        V3 F, T1, T2;
        CInfoOut ci;
        const bool want_ci = REC && B.cinfo != nullptr;
        contact_force<HIST, ROLL>(P, P.comp[2], b1, b2, g, h, F, T1, T2, want_ci ? &ci : nullptr);
        if (want_ci)
            store_cinfo(B.cinfo + (hi + s) * kCInfo, ci.fn, ci.ft, ci.tr, ci.vrot, ci.tc);
        out.F = out.F + F;
        out.T = out.T + T2;
        if (HIST) {
            hcol[hi] = make_double4(h.disp.x, h.disp.y, h.disp.z, steps);
            if (rcol)
                rcol[hi] = h.relvel0;
            out.mask |= 1ull << slot;
        }
        // wrench on the mesh: -F at pt1; T1 is the torque about the body origin
        double wv[6] = {-F.x, -F.y, -F.z, T1.x, T1.y, T1.z};
        const unsigned peers = __match_any_sync(__activemask(), m);
        const unsigned lane = threadIdx.x & 31;
        const unsigned leader = (unsigned)__ffs(peers) - 1u;
        for (unsigned rem = peers & ~(1u << leader); rem; rem &= rem - 1) {
            const int srcl = __ffs(rem) - 1;
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const double o = __shfl_sync(peers, wv[c], srcl);
                if (lane == leader)
                    wv[c] += o;
            }
        }
        if (lane == leader) {
#pragma unroll
            for (int c = 0; c < 6; c++)
                atomicAdd(&MS.wrench[m][c], wv[c]);
        }
    }
}

// --------------------------------------------------------------------------------------------
// fused narrowphase + force + integrate.  One thread per sphere in storage (cell) order.
// --------------------------------------------------------------------------------------------
// One warp per block: there is no block-level cooperation in this kernel, and single-warp blocks free their registers and
// shared memory the moment the warp's longest contact list is done (r01x/r01y: 128 threads 328 us, 64: 316, 32: 311).
#ifndef DEMB200_FORCE_THREADS
#define DEMB200_FORCE_THREADS 32
#endif
constexpr int kForceThreads = DEMB200_FORCE_THREADS;
#ifndef DEMB200_P1_BATCH
#define DEMB200_P1_BATCH 4
#endif
// The streams a block starts with (own position / velocity record, candidate count, first candidate ids) are requested into L2 by
// the block that runs DEMB200_PF_AHEAD blocks earlier (about half of what is resident on the GPU): the first round trip of a warp
// is an L2 hit instead of a DRAM access.  600 / 1200 / 2400: 270.5 / 269.3 / 272.2 us against 272.5 (r02pfa).  0 = off.
#ifndef DEMB200_PF_AHEAD
#define DEMB200_PF_AHEAD 1200
#endif
#ifndef DEMB200_EARLYPF
#define DEMB200_EARLYPF 0  /* prefetch of history rows (1, 2) and partner records (1) at first sight: a loss since the 256-bit gathers (r01v) */
#endif
#ifndef DEMB200_PF2
#define DEMB200_PF2 0
#endif
// bounding experiments (wrong physics, same arithmetic; never the shipped build; results in profiles/README.md): 1 = history rows
// addressed by contact index (coalesced), 2 = no history traffic, 8 / 9 = one more read-modify-write per contact on a scratch row
// (by contact index / by slot)
#ifndef DEMB200_DIAG
#define DEMB200_DIAG 0
#endif
#define DEMB200_DIAG_HROW(k_, slot_) ((DEMB200_DIAG == 1) ? (size_t)(k_) : (size_t)((slot_) & 63u))
// 6 = body 2 of a pair does not write its history copy; 7 = 6 + body 2 reads a row of the PARTNER's column (the access pattern of a
// single history copy kept by body 1)
#define DEMB200_DIAG_HCOL(j_, hi_flag_) ((DEMB200_DIAG == 7 && !(hi_flag_)) ? B.hist + (j_) : hcol)
// the rolling / spinning instantiations carry more live state: 12 warps per SM at 168 registers (no spills) beat 16 warps
// at 128 registers with 172 B of spills (350 vs 400 us on the polydisperse + rolling-friction variant of the bench)
#ifndef DEMB200_ROLL_MINBLOCKS
#define DEMB200_ROLL_MINBLOCKS (384 / DEMB200_FORCE_THREADS)
#endif
#ifndef DEMB200_FORCE_MINBLOCKS
#define DEMB200_FORCE_MINBLOCKS (512 / DEMB200_FORCE_THREADS)  /* 16 warps per SM: 128 registers per thread */
#endif

// rows of the force kernel's shared-memory contact list: K with MultiStep history (more contacts than history slots is an
// error in any case), else the upper bound of K
__host__ __device__ __forceinline__ int force_list_slots(const Params& P, bool hist) {
    return hist ? (P.K < kMaxSlots ? P.K : kMaxSlots) : kMaxSlots;
}
__host__ __device__ __forceinline__ size_t force_smem_bytes(const Params& P, bool hist) {
    return (size_t)force_list_slots(P, hist) * kForceThreads * 5u;
}

// FAST: 0 = generic law (contact_force), 1 = Hertz with material properties, 2 = Hertz with user coefficients
template <bool HIST, bool ROLL, int FAST, bool REC, bool MESH>
#ifdef DEMB200_FORCE_MAXNREG  /* experiment: registers per thread set directly (17 / 18 one-warp blocks per SM = 120 / 112) */
__global__ void __launch_bounds__(kForceThreads) __maxnreg__(ROLL ? 168 : DEMB200_FORCE_MAXNREG) k_force_integrate(
#else
__global__ void __launch_bounds__(kForceThreads, ROLL ? DEMB200_ROLL_MINBLOCKS : DEMB200_FORCE_MINBLOCKS) k_force_integrate(
#endif
    const __grid_constant__ Params P,
                                                                      const __grid_constant__ Buffers B, const unsigned pass) {
    // contact list of the block, dynamic shared memory sized by the launch (force_list_slots): with MultiStep history a sphere may
    // not have more than K contacts anyway, and 16 rows instead of 32 let the driver pick the 64 KB carve-out (192 KB of L1)
    extern __shared__ unsigned force_smem[];
    const int cap = force_list_slots(P, HIST);
    unsigned* const clist = force_smem;                                                        // storage slot of the k-th touching candidate
    unsigned char* const cslot = reinterpret_cast<unsigned char*>(clist + cap * kForceThreads);  // its index in the candidate list (= history slot)
    Ctrl& C = *B.ctrl;
    const unsigned src = C.f_src, dst = src ^ 1u;
    const double4* __restrict__ pos_in = B.pos[src];
    const VelRec* __restrict__ vel_in = B.vel[src];
    const unsigned tid = threadIdx.x;
    // pass 0: every sphere.  Slab mode with the direct halo splits the step in two: pass 1 = every sphere without a ghost among
    // its candidates (runs while the halo is still in flight), pass 2 = the others (Buffers::bnd_list), after the halo has landed.
    unsigned s = blockIdx.x * kForceThreads + tid;
#if DEMB200_PF_AHEAD
    if (pass != 2u) {
        const unsigned sa = s + (unsigned)DEMB200_PF_AHEAD * kForceThreads;
        if (sa < P.N) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pos_in + sa));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vel_in + sa));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(vel_in + sa) + 32));
            if ((tid & 7u) == 0u) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(B.ncnt + sa));
#pragma unroll
                for (int u = 0; u < DEMB200_P1_BATCH; u++)
                    if (u < P.Kn)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(B.nl + (size_t)u * P.Np + sa));
            }
        }
    }
#endif
    if (pass == 2u)
        s = (s < C.n_bnd) ? B.bnd_list[s] : P.N;
    bool valid = s < P.N;
    if (pass == 1u && valid && (B.ncnt[s] & kBndFlag))
        valid = false;
    const GridDev& G = C.mc;

    double4 me = make_double4(0, 0, 0, 1.0);
    VelVal mv;
    mv.v = mk(0, 0, 0); mv.w = mk(0, 0, 0); mv.sid = 0; mv.meta = 0; mv.amask = 0ull;
    int cnt = 0;
    unsigned wcand = 0;  // walls this sphere can reach before the next rebuild (k_build_list)
    unsigned tfirst = 0, tcand = 0;  // mesh triangles in reach: candidate slots tfirst .. tfirst + tcand - 1
    if (valid) {
        me = ld256(pos_in + s);
        mv = load_vel(vel_in, s);
        // ---- phase 1: exact sphere_sphere test (ChNarrowphasePRIMS.cpp:50-59, separation = 0) on the candidates
        if (HIST && DEMB200_EARLYPF) {
            // the live history records of this sphere are a pure stream (read once, by this thread): start them now
            unsigned long long m = mv.amask;
            const double4* hp = B.hist + s;
            while (m) {
                const int k = __ffsll((long long)m) - 1;
                m &= m - 1;
                prefetch_l1(hp + (size_t)k * P.Np);
            }
        }
        const uint32_t* __restrict__ nl = B.nl + s;
        // batches of kP1 candidates: the ids of the next batch and the positions of this batch are in flight together.  The first
        // batch of ids does not wait for the candidate count (rows beyond it hold stale but readable entries, discarded below):
        // count, ids, own position and own velocity record are one round trip instead of two
        constexpr int kP1 = DEMB200_P1_BATCH;
        unsigned jn[kP1];
#pragma unroll
        for (int u = 0; u < kP1; u++)
            jn[u] = (u < P.Kn) ? ld_nl(nl + (size_t)u * P.Np) : 0u;
        const unsigned ncw = B.ncnt[s];
        const unsigned nc = ncw & 0x7Fu;
        wcand = (ncw >> 8) & 0xFFFFu;
        if (MESH) {
            tfirst = nc;
            tcand = ncw >> 24;
        }
#pragma unroll
        for (int u = 0; u < kP1; u++)
            jn[u] = ((unsigned)u < nc) ? jn[u] : s;
        for (unsigned k0 = 0; k0 < nc; k0 += kP1) {
            unsigned jj[kP1];
            double4 pp[kP1];
#pragma unroll
            for (int u = 0; u < kP1; u++)
                jj[u] = jn[u];
#pragma unroll
            for (int u = 0; u < kP1; u++)
                pp[u] = ld256g(pos_in + (jj[u] & ~kEntryFlags));
#pragma unroll
            for (int u = 0; u < kP1; u++)
                jn[u] = (k0 + kP1 + u < nc) ? ld_nl(nl + (size_t)(k0 + kP1 + u) * P.Np) : s;
#pragma unroll
            for (int u = 0; u < kP1; u++) {
                const V3 d = mk(__dsub_rn(pp[u].x, me.x), __dsub_rn(pp[u].y, me.y), __dsub_rn(pp[u].z, me.z));
                const double dist2 = dot_rn(d, d);
                const double rs = __dadd_rn(me.w, pp[u].w);
                // (a padding lane compares the sphere with itself: dist2 = 0 < 1e-12, rejected)
                if (dist2 >= __dmul_rn(rs, rs) || dist2 < 1e-12)
                    continue;
                if (cnt < cap) {
                    clist[cnt * kForceThreads + tid] = jj[u] & ~kEntryFlags;
                    cslot[cnt * kForceThreads + tid] = (unsigned char)((k0 + u) | ((jj[u] & kHiFlag) ? kSlotHi : 0u));
                }
#if DEMB200_EARLYPF == 1
                prefetch_l1(vel_in + (jj[u] & ~kEntryFlags));  // the partner's velocity record is needed in phase 2
#endif
                cnt++;
            }
        }
        if (cnt > cap) {
            atomicOr(&C.err, ERR_HISTORY_OVERFLOW);
            cnt = cap;
        }
    }
    const unsigned sid = mv.sid;
    const unsigned flags = mv.meta & 0xFFu;
    const bool ghost = (flags & FLAG_GHOST) != 0;
    if (ghost)
        cnt = 0;  // a ghost is somebody else's sphere: no forces here, its state arrives with the next halo message
    const unsigned wmask_old = (mv.meta >> 8) & 0xFFFFu;
    const unsigned long long amask_old = mv.amask;
    unsigned wmask_new = 0u;
    unsigned long long amask_new = 0ull;
    const double my_mass0 = sphere_mass(P, me.w);
    V3 Fsum = mk(0, 0, 0), Tsum = mk(0, 0, 0);
    unsigned ncontacts = 0;
    double4* const hcol = HIST ? B.hist + s : nullptr;            // my column of history records
    double* const rcol = (HIST && B.hrel) ? B.hrel + s : nullptr;
    const V3 mpos0 = mk(me.x, me.y, me.z);

    if (valid && !ghost) {
        // ---- walls first: body 1 = wall body (lower id), body 2 = this sphere
        double amin[3], amax[3];
        if (wcand)
            sphere_aabb_offset(me, G.origin, amin, amax);
        while (wcand) {
            const int w = __ffs(wcand) - 1;
            wcand &= wcand - 1;
            const Wall& W = B.walls->w[w];
            if (!W.enabled)
                continue;
            Geom g;
            bool hit;
            if (W.type == WALL_ZCYL) {
                hit = zcyl_sphere_dev(W, mpos0, me.w, g);
            } else if (W.type == WALL_SPHERE) {
                hit = ball_sphere_dev(W, mpos0, me.w, g);
            } else if (W.type == WALL_ZCONE) {
                hit = zcone_sphere_dev(W, mpos0, me.w, g);
            } else if (W.type == WALL_BOX) {
                // broadphase AABB overlap on origin-offset boxes (ChCollisionUtils.h:83-87)
                if (!(amin[0] <= G.wmax[w][0] && G.wmin[w][0] <= amax[0] && amin[1] <= G.wmax[w][1] &&
                      G.wmin[w][1] <= amax[1] && amin[2] <= G.wmax[w][2] && G.wmin[w][2] <= amax[2]))
                    continue;
                hit = box_sphere_dev(W, mpos0, me.w, g);
            } else {
                hit = plane_sphere_dev(W, mpos0, me.w, g);
            }
            if (!hit)
                continue;
            ncontacts++;
            if (REC) {
                unsigned long long at = atomicAdd(&C.pair_count, 1ull);
                if (at < B.pair_cap)
                    B.pairs[at] = ((unsigned long long)w << 32) | (unsigned long long)(P.shape_base + sid);
            }
            if (g.depth >= 0)  // ChIterativeSolverMulticoreSMC.cpp:96-104: no force, no history
                continue;
            Hist h{mk(0, 0, 0), 0.0, 0.0, true};
            double steps = 0.0;
            const size_t hi = (size_t)(P.Kn + w) * P.Np;
            if (HIST && ((wmask_old >> w) & 1u)) {
                const double4 r = ld256v(hcol + hi);
                h.disp = mk(r.x, r.y, r.z);
                steps = r.w;
                h.dur = steps * P.dt;
                if (rcol)
                    h.relvel0 = rcol[hi];
                h.isnew = false;
                steps += 1.0;
            }
            Body b1{mk(0, 0, 0), mk(W.vel[0], W.vel[1], W.vel[2]), mk(0, 0, 0), P.wall_mass};
            if (W.omg[0] != 0.0 || W.omg[1] != 0.0 || W.omg[2] != 0.0) {
                // SetBCPlaneRotation: the wall's material point at the contact moves with vel + omg x (pt1 - rc)
                b1.pos = mk(W.rc[0], W.rc[1], W.rc[2]);
                b1.w = mk(W.omg[0], W.omg[1], W.omg[2]);
            }
            Body b2{mpos0, mv.v, mv.w, my_mass0};
            V3 F, T1, T2;
            CInfoOut ci;
            const bool want_ci = REC && B.cinfo != nullptr;
            contact_force<HIST, ROLL>(P, P.comp[1], b1, b2, g, h, F, T1, T2, want_ci ? &ci : nullptr);
            if (want_ci)
                store_cinfo(B.cinfo + (hi + s) * kCInfo, ci.fn, ci.ft, ci.tr, ci.vrot, ci.tc);
            Fsum = Fsum + F;
            Tsum = Tsum + T2;
            if (P.track_wall_forces) {
                atomicAdd(&C.wall_force[w][0], -F.x);
                atomicAdd(&C.wall_force[w][1], -F.y);
                atomicAdd(&C.wall_force[w][2], -F.z);
            }
            if (HIST) {
                st256h(hcol + hi, make_double4(h.disp.x, h.disp.y, h.disp.z, steps));
                if (rcol)
                    rcol[hi] = h.relvel0;
                wmask_new |= 1u << w;
            }
        }
    }

    if (MESH && tcand && !ghost) {
        MeshOut mo;
        mesh_contacts<HIST, ROLL, REC>(P, B, s, sid, me, mv.v, mv.w, my_mass0, tfirst, tcand, amask_old, mo);
        Fsum = Fsum + mo.F;
        Tsum = Tsum + mo.T;
        amask_new |= mo.mask;
        ncontacts += mo.n;
    }

    // ---- phase 2: all lanes evaluate their k-th sphere contact together (contacts are in stable-id order).
    //      Software pipeline: partner state and history record of contact k+1 are in flight while k is evaluated.
    const int maxc = __reduce_max_sync(0xffffffffu, cnt);
    double4 pj_next = me;
    VelVal ov_next = mv;
    double4 hr_next = make_double4(0, 0, 0, 0);
    unsigned slot_next = 0;
    if (cnt > 0) {
        const unsigned j0 = clist[tid];
        slot_next = cslot[tid];  // slot | kSlotHi
        pj_next = ld256g(pos_in + j0);
        ov_next = load_vel_g(vel_in, j0);
        if (HIST && DEMB200_DIAG != 2 && ((amask_old >> (slot_next & 63u)) & 1ull))
            hr_next = ld256v(DEMB200_DIAG_HCOL(j0, slot_next & kSlotHi) + DEMB200_DIAG_HROW(0, slot_next) * P.Np);
    }
    for (int k = 0; k < maxc; k++) {
        if (k >= cnt)
            continue;
        const double4 pj = pj_next;
        const VelVal ov = ov_next;
        double4 hr = hr_next;
        // Keep the consumer of the prefetched record HERE: without this the compiler copies the freshly loaded
        // registers into their loop-carried homes right behind the load below and stalls on it (ncu r01d).
#ifndef DEMB200_NO_HRPIN
        asm volatile("" : "+d"(hr.x), "+d"(hr.y), "+d"(hr.z), "+d"(hr.w));
#endif
        const unsigned slot = slot_next & 63u;
        // canonical orientation: body 1 = lower stable id.  Decided by k_build_list (kHiFlag) and carried in the slot
        // byte: the partner's id is not gathered for it (the only other use of that id is the recorded pair key).
        const bool me1 = (slot_next & kSlotHi) != 0;
#if DEMB200_PF2
        if (k + 2 < cnt) {  // two contacts ahead: lines into L1, so that the register loads below hit
            const unsigned j2 = clist[(k + 2) * kForceThreads + tid];
            const unsigned s2 = cslot[(k + 2) * kForceThreads + tid] & 63u;
            prefetch_l1(pos_in + j2);
            prefetch_l1(vel_in + j2);
            if (HIST && ((amask_old >> s2) & 1ull))
                prefetch_l1(hcol + (size_t)s2 * P.Np);
        }
#endif
        if (k + 1 < cnt) {
            const unsigned jn = clist[(k + 1) * kForceThreads + tid];
            slot_next = cslot[(k + 1) * kForceThreads + tid];
            pj_next = ld256g(pos_in + jn);
            ov_next = load_vel_g(vel_in, jn);
            if (HIST && DEMB200_DIAG != 2 && ((amask_old >> (slot_next & 63u)) & 1ull))
                hr_next = ld256v(DEMB200_DIAG_HCOL(jn, slot_next & kSlotHi) + DEMB200_DIAG_HROW(k + 1, slot_next) * P.Np);
        }
        const unsigned sj = REC ? ov.sid : 0u;
        const bool had = HIST && ((amask_old >> slot) & 1ull);
        const size_t hi = DEMB200_DIAG_HROW(k, slot) * P.Np;
        ncontacts++;
        if (REC && !me1) {
            unsigned long long at = atomicAdd(&C.pair_count, 1ull);
            if (at < B.pair_cap)
                B.pairs[at] = ((unsigned long long)(P.shape_base + sj) << 32) | (unsigned long long)(P.shape_base + sid);
        }
        if (FAST) {
            // sphere_sphere geometry in my own frame (n from me to the partner); |delta|^2 as in the contact test
            const V3 delta = mk(__dsub_rn(pj.x, me.x), __dsub_rn(pj.y, me.y), __dsub_rn(pj.z, me.z));
            const double d2 = dot_rn(delta, delta);
            const double inv_d = fast_rsqrt(d2);
            const double dist = d2 * inv_d;
            if (dist - __dadd_rn(me.w, pj.w) >= 0)
                continue;
            const V3 n = delta * inv_d;
            V3 disp = mk(hr.x, hr.y, hr.z);
            double steps = hr.w;
            if (had && !me1)
                disp = -disp;  // canonical (body 1 -> body 2) to my frame
            V3 F, T;
            CInfoOut ci;
            const bool want_ci = REC && B.cinfo != nullptr;
            sphere_contact_fast<HIST, ROLL, FAST == 1>(P, P.comp[0], n, dist, me.w, pj.w, mv.v, mv.w, ov.v, ov.w, my_mass0,
                                            sphere_mass(P, pj.w), me1, disp, steps, !had, F, T, want_ci ? &ci : nullptr);
            if (want_ci)
                store_cinfo(B.cinfo + (hi + s) * kCInfo, ci.fn, ci.ft, ci.tr, ci.vrot, ci.tc);
            Fsum = Fsum + F;
            Tsum = Tsum + T;
            if (HIST) {
                if (!me1)
                    disp = -disp;
                if (DEMB200_DIAG != 2 && !((DEMB200_DIAG == 6 || DEMB200_DIAG == 7) && !me1))
                    st256h(hcol + hi, make_double4(disp.x, disp.y, disp.z, steps));
#if DEMB200_DIAG == 8 || DEMB200_DIAG == 9
                {   // marginal cost of one more 32-byte read + write per contact on a scratch array (the rebuild's staging rows):
                    // 8 = addressed by contact index (coalesced across the warp), 9 = by candidate slot (scattered like the history)
                    double4* const q = B.stage + s + (size_t)(DEMB200_DIAG == 8 ? (unsigned)k & 15u : slot & 15u) * P.Np;
                    double4 t = ld256v(q);
                    t.x += disp.x;
                    st256h(q, t);
                }
#endif
                amask_new |= 1ull << slot;
            }
        } else {
            Body b1, b2;
            double r1, r2;
            {
                Body bm{mpos0, mv.v, mv.w, my_mass0};
                Body bo{mk(pj.x, pj.y, pj.z), ov.v, ov.w, sphere_mass(P, pj.w)};
                b1 = me1 ? bm : bo;
                b2 = me1 ? bo : bm;
                r1 = me1 ? me.w : pj.w;
                r2 = me1 ? pj.w : me.w;
            }
            // sphere_sphere contact geometry, ChNarrowphasePRIMS.cpp:61-69
            Geom g;
            {
                const V3 delta = mk(__dsub_rn(b2.pos.x, b1.pos.x), __dsub_rn(b2.pos.y, b1.pos.y), __dsub_rn(b2.pos.z, b1.pos.z));
                const double dist = sqrt(dot_rn(delta, delta));
                g.n = mk(__ddiv_rn(delta.x, dist), __ddiv_rn(delta.y, dist), __ddiv_rn(delta.z, dist));
                g.pt1 = mk(__dadd_rn(b1.pos.x, __dmul_rn(g.n.x, r1)), __dadd_rn(b1.pos.y, __dmul_rn(g.n.y, r1)),
                           __dadd_rn(b1.pos.z, __dmul_rn(g.n.z, r1)));
                g.pt2 = mk(__dsub_rn(b2.pos.x, __dmul_rn(g.n.x, r2)), __dsub_rn(b2.pos.y, __dmul_rn(g.n.y, r2)),
                           __dsub_rn(b2.pos.z, __dmul_rn(g.n.z, r2)));
                const double radSum = __dadd_rn(r1, r2);
                g.depth = __dsub_rn(dist, radSum);
                g.erad = __ddiv_rn(__dmul_rn(r1, r2), radSum);
            }
            if (g.depth >= 0)  // ChIterativeSolverMulticoreSMC.cpp:96-104: no force, no history
                continue;
            Hist h{mk(0, 0, 0), 0.0, 0.0, true};
            double steps = 0.0;
            if (had) {
                h.disp = mk(hr.x, hr.y, hr.z);
                steps = hr.w;
                h.dur = steps * P.dt;
                if (rcol)
                    h.relvel0 = rcol[hi];
                h.isnew = false;
                steps += 1.0;
            }
            V3 F, T1, T2;
            CInfoOut ci;
            const bool want_ci = REC && B.cinfo != nullptr;
            contact_force<HIST, ROLL>(P, P.comp[0], b1, b2, g, h, F, T1, T2, want_ci ? &ci : nullptr);
            if (want_ci) {  // the generic law reports body 2's view
                if (me1)
                    store_cinfo(B.cinfo + (hi + s) * kCInfo, -ci.fn, -ci.ft, ci.tr1, ci.vrot, ci.tc);
                else
                    store_cinfo(B.cinfo + (hi + s) * kCInfo, ci.fn, ci.ft, ci.tr, ci.vrot, ci.tc);
            }
            if (me1) {
                Fsum = Fsum - F;
                Tsum = Tsum + T1;
            } else {
                Fsum = Fsum + F;
                Tsum = Tsum + T2;
            }
            if (HIST) {
                st256h(hcol + hi, make_double4(h.disp.x, h.disp.y, h.disp.z, steps));
                if (rcol)
                    rcol[hi] = h.relvel0;
                amask_new |= 1ull << slot;
            }
        }
    }

    double nmnx = CUDART_INF, nmny = CUDART_INF, nmnz = CUDART_INF, nmxx = -CUDART_INF, nmxy = -CUDART_INF, nmxz = -CUDART_INF;
    double dx2 = 0.0;
    // pass 1 runs next to the halo pick-up (side stream), which rewrites the ghost records of BOTH buffers itself: ghosts sit out
    if (valid && !(pass == 1u && ghost)) {
        if (HIST && __popcll(amask_new) + __popc(wmask_new) > P.K)
            atomicOr(&C.err, ERR_HISTORY_OVERFLOW);
        if (REC) {
            B.recF[3 * (size_t)sid + 0] = Fsum.x; B.recF[3 * (size_t)sid + 1] = Fsum.y; B.recF[3 * (size_t)sid + 2] = Fsum.z;
            B.recT[3 * (size_t)sid + 0] = Tsum.x; B.recT[3 * (size_t)sid + 1] = Tsum.y; B.recT[3 * (size_t)sid + 2] = Tsum.z;
            atomicAdd(&C.n_contacts, (unsigned long long)ncontacts);
        }

        // ---- time integration ----
        const bool fixed = (flags & (FLAG_FIXED | FLAG_GHOST)) != 0;
        // own state re-derived from `me` here (the lean phase 2 parks it in shared memory and reloads it: nothing of it stays
        // live in registers across the contact loop)
        const V3 mpos = mk(me.x, me.y, me.z);
        const double my_mass = sphere_mass(P, me.w);
        const double hdt = P.dt;
        const double inv_m = 1.0 / my_mass;
        const double inv_I = 1.0 / (0.4 * my_mass * me.w * me.w);
        const V3 gv = mk(P.g[0], P.g[1], P.g[2]);
        V3 x = mpos;
        V3 vn = mv.v, wn = mv.w;
        if (!fixed) {
            if (P.integrator == 2) {
                // Multicore: hf = h*(g*m) + h*F; v+ = v + M^-1 hf; x+ = x + v+ h (ChBody.cpp:247-256,288-297)
                V3 hf = hdt * (gv * my_mass) + hdt * Fsum;
                vn = mv.v + inv_m * hf;
                wn = mv.w + inv_I * (hdt * Tsum);
                x = x + vn * hdt;
            } else {
                const V3 acc = gv + inv_m * Fsum;
                const V3 alp = inv_I * Tsum;
                if (P.integrator == 3) {         // extended Taylor (ChDemSMC.cuh:1347-1351)
                    x = x + hdt * (mv.v + 0.5 * hdt * acc);
                    vn = mv.v + hdt * acc;
                    wn = mv.w + hdt * alp;
                } else if (P.integrator == 0) {  // forward Euler
                    x = x + hdt * mv.v;
                    vn = mv.v + hdt * acc;
                    wn = mv.w + hdt * alp;
                } else {                         // Chung (ChDemSMC.cuh:1266-1277): beta = 28/27, gamma = 3/2
                    const double2* ap = reinterpret_cast<const double2*>(B.acc[src] + 6 * (size_t)s);
                    const double2 o0 = ap[0], o1 = ap[1], o2 = ap[2];
                    const V3 ao = mk(o0.x, o0.y, o1.x), lo = mk(o1.y, o2.x, o2.y);
                    const double beta = 28.0 / 27.0;
                    x = x + hdt * (mv.v + hdt * (beta * acc + (0.5 - beta) * ao));
                    vn = mv.v + hdt * (1.5 * acc - 0.5 * ao);
                    wn = mv.w + hdt * (1.5 * alp - 0.5 * lo);
                    double2* aw = reinterpret_cast<double2*>(B.acc[dst] + 6 * (size_t)s);
                    aw[0] = make_double2(acc.x, acc.y);
                    aw[1] = make_double2(acc.z, alp.x);
                    aw[2] = make_double2(alp.y, alp.z);
                }
            }
        } else if (P.integrator == 1) {
            double2* aw = reinterpret_cast<double2*>(B.acc[dst] + 6 * (size_t)s);
            aw[0] = aw[1] = aw[2] = make_double2(0.0, 0.0);
        }
        if (!(isfinite(x.x) && isfinite(x.y) && isfinite(x.z)))
            atomicOr(&C.err, ERR_NAN);
        st256s(B.pos[dst] + s, make_double4(x.x, x.y, x.z, me.w));
        store_vel_s(B.vel[dst], s, vn, wn, sid, flags | (wmask_new << 8), amask_new);
        if (!ghost) {
            nmnx = x.x - me.w; nmny = x.y - me.w; nmnz = x.z - me.w;
            nmxx = x.x + me.w; nmxy = x.y + me.w; nmxz = x.z + me.w;
        }
        const V3 dxv = x - mpos;
        dx2 = dot(dxv, dxv);
    }
    // running bounding box of the new sphere AABBs -> next step's grids; largest displacement -> Verlet travel.
    // The accumulator starts at the walls' box: a warp whose spheres all lie inside what is already recorded (every warp
    // of a bed inside its container) skips the reduction.
    {
        const bool grows = nmnx < dec_ord(C.bbox[0]) || nmny < dec_ord(C.bbox[1]) || nmnz < dec_ord(C.bbox[2]) ||
                           nmxx > dec_ord(C.bbox[3]) || nmxy > dec_ord(C.bbox[4]) || nmxz > dec_ord(C.bbox[5]);
        if (__any_sync(0xffffffffu, grows))
            block_bbox_commit(nmnx, nmny, nmnz, nmxx, nmxy, nmxz, C.bbox);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        dx2 = fmax(dx2, __shfl_xor_sync(0xffffffffu, dx2, o));
    if ((tid & 31) == 0) {
        const unsigned long long e = (unsigned long long)__double_as_longlong(dx2);  // non-negative: raw bits are ordered
        if (e > C.max_dx2)
            atomicMax(&C.max_dx2, e);
    }
}

// --------------------------------------------------------------------------------------------
// Tile force kernel: the same step as k_force_integrate, with the neighbour bins staged in shared memory.
//
// One block per tile of the tiled search grid (kTile^3 cells, Params::tiled).  The spheres of the tile are a contiguous run
// of the storage order; every candidate of every one of them lies in the tile's box (the tile plus a one-cell halo,
// kBox^3 cells).  The block copies position, radius, velocity and angular velocity of all spheres of the box into shared
// memory once (coalesced runs, 80 bytes per sphere); after that the exact narrowphase on the candidates (phase 1) and the
// partner state of every contact (phase 2) are shared-memory reads addressed by the 16-bit candidate code written at list
// build time (box cell, rank in cell), instead of three divergent global gathers per contact.  Only the history rows (own
// column, read and written once) and the sphere's own record still go to global memory.  Arithmetic, contact order and
// therefore every bit of the result are those of k_force_integrate (checked by the parity tests, which run both).
// Restrictions: Hertz fast paths (FAST 1 / 2), no meshes, no recording -- everything else runs k_force_integrate.
// --------------------------------------------------------------------------------------------
#ifndef DEMB200_TILE_THREADS
#define DEMB200_TILE_THREADS 160
#endif
#ifndef DEMB200_TILE_CAP
#define DEMB200_TILE_CAP 640  /* sphere records staged per block; a box holds ~430 - 600 of them in a dense bed */
#endif
#ifndef DEMB200_TILE_MINBLOCKS
#define DEMB200_TILE_MINBLOCKS 3
#endif
constexpr int kTileThreads = DEMB200_TILE_THREADS;
constexpr int kTileCap = DEMB200_TILE_CAP;
struct __align__(16) TileRec {
    double x, y, z, r, vx, vy, vz, wx, wy, wz;
};
constexpr size_t kTileSmemBytes = (size_t)kTileCap * sizeof(TileRec) + (size_t)kMaxSlots * kTileThreads * 3;

// walls of one sphere: body 1 = wall body (lower id), body 2 = the sphere (same code path as in k_force_integrate)
template <bool HIST, bool ROLL>
__device__ __forceinline__ void wall_contacts(const Params& P, const Buffers& B, Ctrl& C, const double4 me, const V3 v, const V3 w,
                                              unsigned wcand, const unsigned wmask_old, double4* const hcol, double* const rcol,
                                              const double my_mass, V3& Fsum, V3& Tsum, unsigned& wmask_new) {
    const GridDev& G = C.mc;
    const V3 mpos = mk(me.x, me.y, me.z);
    double amin[3], amax[3];
    if (wcand)
        sphere_aabb_offset(me, G.origin, amin, amax);
    while (wcand) {
        const int wi = __ffs(wcand) - 1;
        wcand &= wcand - 1;
        const Wall& W = B.walls->w[wi];
        if (!W.enabled)
            continue;
        Geom g;
        bool hit;
        if (W.type == WALL_ZCYL) {
            hit = zcyl_sphere_dev(W, mpos, me.w, g);
        } else if (W.type == WALL_SPHERE) {
            hit = ball_sphere_dev(W, mpos, me.w, g);
        } else if (W.type == WALL_ZCONE) {
            hit = zcone_sphere_dev(W, mpos, me.w, g);
        } else if (W.type == WALL_BOX) {
            if (!(amin[0] <= G.wmax[wi][0] && G.wmin[wi][0] <= amax[0] && amin[1] <= G.wmax[wi][1] &&
                  G.wmin[wi][1] <= amax[1] && amin[2] <= G.wmax[wi][2] && G.wmin[wi][2] <= amax[2]))
                continue;
            hit = box_sphere_dev(W, mpos, me.w, g);
        } else {
            hit = plane_sphere_dev(W, mpos, me.w, g);
        }
        if (!hit || g.depth >= 0)
            continue;
        Hist h{mk(0, 0, 0), 0.0, 0.0, true};
        double steps = 0.0;
        const size_t hi = (size_t)(P.Kn + wi) * P.Np;
        if (HIST && ((wmask_old >> wi) & 1u)) {
            const double4 r = ld256v(hcol + hi);
            h.disp = mk(r.x, r.y, r.z);
            steps = r.w;
            h.dur = steps * P.dt;
            if (rcol)
                h.relvel0 = rcol[hi];
            h.isnew = false;
            steps += 1.0;
        }
        Body b1{mk(0, 0, 0), mk(W.vel[0], W.vel[1], W.vel[2]), mk(0, 0, 0), P.wall_mass};
        if (W.omg[0] != 0.0 || W.omg[1] != 0.0 || W.omg[2] != 0.0) {
            b1.pos = mk(W.rc[0], W.rc[1], W.rc[2]);
            b1.w = mk(W.omg[0], W.omg[1], W.omg[2]);
        }
        Body b2{mpos, v, w, my_mass};
        V3 F, T1, T2;
        contact_force<HIST, ROLL>(P, P.comp[1], b1, b2, g, h, F, T1, T2);
        Fsum = Fsum + F;
        Tsum = Tsum + T2;
        if (P.track_wall_forces) {
            atomicAdd(&C.wall_force[wi][0], -F.x);
            atomicAdd(&C.wall_force[wi][1], -F.y);
            atomicAdd(&C.wall_force[wi][2], -F.z);
        }
        if (HIST) {
            st256(hcol + hi, make_double4(h.disp.x, h.disp.y, h.disp.z, steps));
            if (rcol)
                rcol[hi] = h.relvel0;
            wmask_new |= 1u << wi;
        }
    }
}

// time integration of one sphere and the store of its new record (same arithmetic as the tail of k_force_integrate)
__device__ __forceinline__ V3 integrate_store(const Params& P, const Buffers& B, Ctrl& C, unsigned src, unsigned dst, unsigned s,
                                              const double4 me, const V3 v0, const V3 w0, unsigned sid, unsigned flags, const V3 Fsum,
                                              const V3 Tsum, unsigned wmask_new, unsigned long long amask_new) {
    const V3 mpos = mk(me.x, me.y, me.z);
    const double my_mass = sphere_mass(P, me.w);
    const bool fixed = (flags & (FLAG_FIXED | FLAG_GHOST)) != 0;
    const double hdt = P.dt;
    const double inv_m = 1.0 / my_mass;
    const double inv_I = 1.0 / (0.4 * my_mass * me.w * me.w);
    const V3 gv = mk(P.g[0], P.g[1], P.g[2]);
    V3 x = mpos;
    V3 vn = v0, wn = w0;
    if (!fixed) {
        if (P.integrator == 2) {
            V3 hf = hdt * (gv * my_mass) + hdt * Fsum;
            vn = v0 + inv_m * hf;
            wn = w0 + inv_I * (hdt * Tsum);
            x = x + vn * hdt;
        } else {
            const V3 acc = gv + inv_m * Fsum;
            const V3 alp = inv_I * Tsum;
            if (P.integrator == 3) {
                x = x + hdt * (v0 + 0.5 * hdt * acc);
                vn = v0 + hdt * acc;
                wn = w0 + hdt * alp;
            } else if (P.integrator == 0) {
                x = x + hdt * v0;
                vn = v0 + hdt * acc;
                wn = w0 + hdt * alp;
            } else {
                const double2* ap = reinterpret_cast<const double2*>(B.acc[src] + 6 * (size_t)s);
                const double2 o0 = ap[0], o1 = ap[1], o2 = ap[2];
                const V3 ao = mk(o0.x, o0.y, o1.x), lo = mk(o1.y, o2.x, o2.y);
                const double beta = 28.0 / 27.0;
                x = x + hdt * (v0 + hdt * (beta * acc + (0.5 - beta) * ao));
                vn = v0 + hdt * (1.5 * acc - 0.5 * ao);
                wn = w0 + hdt * (1.5 * alp - 0.5 * lo);
                double2* aw = reinterpret_cast<double2*>(B.acc[dst] + 6 * (size_t)s);
                aw[0] = make_double2(acc.x, acc.y);
                aw[1] = make_double2(acc.z, alp.x);
                aw[2] = make_double2(alp.y, alp.z);
            }
        }
    } else if (P.integrator == 1) {
        double2* aw = reinterpret_cast<double2*>(B.acc[dst] + 6 * (size_t)s);
        aw[0] = aw[1] = aw[2] = make_double2(0.0, 0.0);
    }
    if (!(isfinite(x.x) && isfinite(x.y) && isfinite(x.z)))
        atomicOr(&C.err, ERR_NAN);
    st256(B.pos[dst] + s, make_double4(x.x, x.y, x.z, me.w));
    store_vel(B.vel[dst], s, vn, wn, sid, flags | (wmask_new << 8), amask_new);
    return x;
}

template <bool HIST, bool ROLL, int FAST>
__global__ void __launch_bounds__(kTileThreads, DEMB200_TILE_MINBLOCKS) k_force_tile(const __grid_constant__ Params P,
                                                                                   const __grid_constant__ Buffers B, const unsigned pass) {
    static_assert(FAST == 1 || FAST == 2, "the tile kernel carries the Hertz fast paths only");
    extern __shared__ __align__(16) unsigned char tile_smem[];
    TileRec* const s_rec = reinterpret_cast<TileRec*>(tile_smem);
    unsigned short* const clist = reinterpret_cast<unsigned short*>(s_rec + kTileCap);           // [kMaxSlots][kTileThreads] code of the k-th touching candidate
    unsigned char* const cslot = reinterpret_cast<unsigned char*>(clist + kMaxSlots * kTileThreads);  // its candidate slot (= history slot)
    __shared__ unsigned s_off[kBoxCells + 1];   // first staged record of every box cell
    __shared__ unsigned s_gstart[kBoxCells];    // its first sphere in the storage order
    Ctrl& C = *B.ctrl;
    const unsigned tid = threadIdx.x;
    const unsigned ntiles = (unsigned)(C.t_dim[0] * C.t_dim[1] * C.t_dim[2]);
    const unsigned tile = blockIdx.x;
    if (tile >= ntiles)
        return;
    const unsigned t_begin = B.cell_start[(size_t)tile * kTileCells], t_end = B.cell_start[(size_t)(tile + 1) * kTileCells];
    if (t_end <= t_begin)
        return;  // empty tile (the whole block leaves)
    if (pass == 2u) {  // only the tiles that hold a sphere with a ghost candidate have work in the second pass
        bool any = false;
        for (unsigned i = t_begin + threadIdx.x; i < t_end; i += kTileThreads)
            any |= (B.ncnt[i] & kBndFlag) != 0u;
        if (!__syncthreads_or(any))
            return;
    }
    const unsigned src = C.f_src, dst = src ^ 1u;
    const double4* __restrict__ pos_in = B.pos[src];
    const VelRec* __restrict__ vel_in = B.vel[src];

    // ---- stage the box: cell ranges, exclusive scan, records
    {
        const int tx = (int)(tile % (unsigned)C.t_dim[0]), ty = (int)((tile / (unsigned)C.t_dim[0]) % (unsigned)C.t_dim[1]),
                  tz = (int)(tile / (unsigned)(C.t_dim[0] * C.t_dim[1]));
        for (unsigned c = tid; c < (unsigned)kBoxCells; c += kTileThreads) {
            const int x = tx * kTile - 1 + (int)(c % kBox), y = ty * kTile - 1 + (int)((c / kBox) % kBox),
                      z = tz * kTile - 1 + (int)(c / (kBox * kBox));
            unsigned gs = 0u, n = 0u;
            if (x >= 0 && y >= 0 && z >= 0 && x < C.s_dim[0] && y < C.s_dim[1] && z < C.s_dim[2]) {
                const unsigned ci = cell_index(P, C, x, y, z);
                gs = B.cell_start[ci];
                n = B.cell_start[ci + 1] - gs;
            }
            s_gstart[c] = gs;
            s_off[c + 1] = n;
        }
        if (tid == 0)
            s_off[0] = 0u;
        __syncthreads();
        if (tid < 32) {  // inclusive scan of s_off[1 .. kBoxCells] by one warp, kBoxCells / 32 (rounded up) entries per lane
            constexpr unsigned per = (kBoxCells + 31) / 32;
            unsigned sum = 0u;
            for (unsigned q = 0; q < per; q++) {
                const unsigned i = tid * per + q;
                if (i < (unsigned)kBoxCells)
                    sum += s_off[i + 1];
            }
            unsigned inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                if (tid >= (unsigned)o)
                    inc += t;
            }
            unsigned run = inc - sum;
            for (unsigned q = 0; q < per; q++) {
                const unsigned i = tid * per + q;
                if (i < (unsigned)kBoxCells) {
                    run += s_off[i + 1];
                    s_off[i + 1] = run;
                }
            }
        }
        __syncthreads();
        for (unsigned c = tid; c < (unsigned)kBoxCells; c += kTileThreads) {
            const unsigned o = s_off[c], gs = s_gstart[c];
            const unsigned n = min(s_off[c + 1] - o, (unsigned)kTileCap > o ? (unsigned)kTileCap - o : 0u);
            constexpr unsigned kS = 4;  // records of one cell in flight together (a cell holds ~2 spheres)
            for (unsigned r0 = 0; r0 < n; r0 += kS) {
                double4 p[kS], q0[kS];
                double2 q1[kS];
#pragma unroll
                for (unsigned u = 0; u < kS; u++)
                    if (r0 + u < n) {
                        p[u] = ld256(pos_in + gs + r0 + u);
                        q0[u] = ld256(vel_in + gs + r0 + u);
                        q1[u] = *reinterpret_cast<const double2*>(reinterpret_cast<const char*>(vel_in + gs + r0 + u) + 32);
                    }
#pragma unroll
                for (unsigned u = 0; u < kS; u++)
                    if (r0 + u < n) {
                        // TileRec = (x y z r | vx vy vz wx | wy wz): 80 bytes, five 16-byte shared stores
                        double2* d = reinterpret_cast<double2*>(s_rec + o + r0 + u);
                        d[0] = make_double2(p[u].x, p[u].y);
                        d[1] = make_double2(p[u].z, p[u].w);
                        d[2] = make_double2(q0[u].x, q0[u].y);
                        d[3] = make_double2(q0[u].z, q0[u].w);
                        d[4] = q1[u];
                    }
            }
        }
    }
    // A box that does not fit the staging buffer, or a cell too crowded for the 6-bit rank of the candidate code, sends the
    // whole block through the instantiation that can fall back to the global arrays.  (Kept out of the common instantiation:
    // a never-taken global load at the head of the contact loop would share its scoreboard with the history prefetch.)
    bool crowded = false;
    for (unsigned c = tid; c < (unsigned)kBoxCells; c += kTileThreads)
        crowded |= (s_off[c + 1] - s_off[c]) >= kCodeNoRank;
    const bool slow_block = __syncthreads_or(crowded || s_off[kBoxCells] > (unsigned)kTileCap) != 0;  // also the barrier after staging

    const unsigned nt = t_end - t_begin;
    auto rounds = [&](auto slow_tag) {
    constexpr bool SLOW = decltype(slow_tag)::value;
    for (unsigned base = 0; base < nt; base += kTileThreads) {
        if (base + (tid & ~31u) >= nt)
            break;  // this warp has no sphere in this round (warp-uniform)
        const unsigned s = t_begin + base + tid;
        bool valid = base + tid < nt;
        if (pass != 0u && valid && (((B.ncnt[s] & kBndFlag) != 0u) != (pass == 2u)))
            valid = false;  // pass 1: spheres without a ghost candidate; pass 2: the others (see k_force_integrate)
        double4 me = make_double4(0, 0, 0, 1.0);
        VelVal mv;
        mv.v = mk(0, 0, 0); mv.w = mk(0, 0, 0); mv.sid = 0; mv.meta = 0; mv.amask = 0ull;
        int cnt = 0;
        unsigned wcand = 0;
        // partner record of a candidate code: staged copy, or (box overflow / rank not encodable) the global arrays
        auto partner_index = [&](unsigned code, unsigned slot, unsigned& idx) -> bool {
            const unsigned c = (code >> kCodeRankBits) & 0xFFu, r = code & kCodeNoRank;
            idx = s_off[c] + r;
            if constexpr (!SLOW) {
                return true;
            } else {
                if (r != kCodeNoRank && idx < (unsigned)kTileCap)
                    return true;
                idx = (r != kCodeNoRank) ? s_gstart[c] + r : (B.nl[(size_t)slot * P.Np + s] & ~(kEntryFlags | kTriFlag));
                return false;
            }
        };
        if (valid) {
            me = ld256(pos_in + s);
            mv = load_vel(vel_in, s);
            const unsigned ncw = B.ncnt[s];
            const unsigned nc = ncw & 0x7Fu;
            wcand = (ncw >> 8) & 0xFFFFu;
            const unsigned short* __restrict__ nl16 = B.nl16 + s;
            // ---- phase 1: exact sphere_sphere test (ChNarrowphasePRIMS.cpp:50-59) on the candidates, positions from the box.
            //      The candidate codes are a coalesced global stream: batches of kP1 of them are in flight together.
            constexpr int kP1 = 8;
            for (unsigned k0 = 0; k0 < nc; k0 += kP1) {
                unsigned codes[kP1];
#pragma unroll
                for (int u = 0; u < kP1; u++)
                    codes[u] = (k0 + u < nc) ? (unsigned)nl16[(size_t)(k0 + u) * P.Np] : 0xFFFFFFFFu;
#pragma unroll
                for (int u = 0; u < kP1; u++) {
                    const unsigned code = codes[u];
                    if (code == 0xFFFFFFFFu)
                        continue;
                    const unsigned k = k0 + u;
                    unsigned idx;
                    double4 pp;
                    if (partner_index(code, k, idx)) {
                        const double2* q = reinterpret_cast<const double2*>(s_rec + idx);
                        const double2 a = q[0], b = q[1];
                        pp = make_double4(a.x, a.y, b.x, b.y);
                    } else {
                        pp = ld256(pos_in + idx);
                    }
                    const V3 d = mk(__dsub_rn(pp.x, me.x), __dsub_rn(pp.y, me.y), __dsub_rn(pp.z, me.z));
                    const double dist2 = dot_rn(d, d);
                    const double rs = __dadd_rn(me.w, pp.w);
                    if (dist2 >= __dmul_rn(rs, rs) || dist2 < 1e-12)
                        continue;
                    if (cnt < kMaxSlots) {
                        clist[cnt * kTileThreads + tid] = (unsigned short)code;
                        cslot[cnt * kTileThreads + tid] = (unsigned char)k;
                    }
                    cnt++;
                }
            }
            if (cnt > kMaxSlots) {
                atomicOr(&C.err, ERR_HISTORY_OVERFLOW);
                cnt = kMaxSlots;
            }
        }
        const unsigned sid = mv.sid;
        const unsigned flags = mv.meta & 0xFFu;
        const bool ghost = (flags & FLAG_GHOST) != 0;
        if (ghost)
            cnt = 0;
        const unsigned wmask_old = (mv.meta >> 8) & 0xFFFFu;
        const unsigned long long amask_old = mv.amask;
        unsigned wmask_new = 0u;
        unsigned long long amask_new = 0ull;
        V3 Fsum = mk(0, 0, 0), Tsum = mk(0, 0, 0);
        double4* const hcol = HIST ? B.hist + s : nullptr;
        double* const rcol = (HIST && B.hrel) ? B.hrel + s : nullptr;
        if (valid && !ghost && wcand)
            wall_contacts<HIST, ROLL>(P, B, C, me, mv.v, mv.w, wcand, wmask_old, hcol, rcol, sphere_mass(P, me.w), Fsum, Tsum, wmask_new);

        // ---- phase 2: all lanes evaluate their k-th sphere contact together (stable-id order); only the history record of the
        //      next contact is prefetched -- the partner state is a shared-memory read
        const int maxc = __reduce_max_sync(0xffffffffu, cnt);
        double4 hr_next = make_double4(0, 0, 0, 0);
        if (HIST && cnt > 0) {
            const unsigned s0 = cslot[tid];
            if ((amask_old >> s0) & 1ull)
                hr_next = ld256v(hcol + (size_t)s0 * P.Np);
        }
        const double my_mass = sphere_mass(P, me.w);
        for (int k = 0; k < maxc; k++) {
            if (k >= cnt)
                continue;
            const unsigned code = clist[k * kTileThreads + tid];
            const unsigned slot = cslot[k * kTileThreads + tid];
            const bool me1 = (code & kCodeHi) != 0;
            double4 hr = hr_next;
            if (HIST && k + 1 < cnt) {
                const unsigned sn = cslot[(k + 1) * kTileThreads + tid];
                if ((amask_old >> sn) & 1ull)
                    hr_next = ld256v(hcol + (size_t)sn * P.Np);
            }
            unsigned idx;
            double px, py, pz, pr;
            V3 vb, wb;
            if (partner_index(code, slot, idx)) {
                const double2* q = reinterpret_cast<const double2*>(s_rec + idx);
                const double2 a = q[0], b = q[1], c2 = q[2], d2_ = q[3], e2 = q[4];
                px = a.x; py = a.y; pz = b.x; pr = b.y;
                vb = mk(c2.x, c2.y, d2_.x);
                wb = mk(d2_.y, e2.x, e2.y);
            } else {
                const double4 pj = ld256(pos_in + idx);
                const double4 q0 = ld256(vel_in + idx);
                const double2 q1 = *reinterpret_cast<const double2*>(reinterpret_cast<const char*>(vel_in + idx) + 32);
                px = pj.x; py = pj.y; pz = pj.z; pr = pj.w;
                vb = mk(q0.x, q0.y, q0.z);
                wb = mk(q0.w, q1.x, q1.y);
            }
            const bool had = HIST && ((amask_old >> slot) & 1ull);
            const size_t hi = (size_t)slot * P.Np;
            const V3 delta = mk(__dsub_rn(px, me.x), __dsub_rn(py, me.y), __dsub_rn(pz, me.z));
            const double d2 = dot_rn(delta, delta);
            const double inv_d = fast_rsqrt(d2);
            const double dist = d2 * inv_d;
            if (dist - __dadd_rn(me.w, pr) >= 0)
                continue;
            const V3 n = delta * inv_d;
            V3 disp = mk(hr.x, hr.y, hr.z);
            double steps = hr.w;
            if (!had) {
                disp = mk(0, 0, 0);
                steps = 0.0;
            } else if (!me1) {
                disp = -disp;  // canonical (body 1 -> body 2) to my frame
            }
            V3 F, T;
            sphere_contact_fast<HIST, ROLL, FAST == 1>(P, P.comp[0], n, dist, me.w, pr, mv.v, mv.w, vb, wb, my_mass, sphere_mass(P, pr), me1,
                                                       disp, steps, !had, F, T);
            Fsum = Fsum + F;
            Tsum = Tsum + T;
            if (HIST) {
                if (!me1)
                    disp = -disp;
                if (DEMB200_DIAG != 2 && !((DEMB200_DIAG == 6 || DEMB200_DIAG == 7) && !me1))
                    st256(hcol + hi, make_double4(disp.x, disp.y, disp.z, steps));
                amask_new |= 1ull << slot;
            }
        }

        double nmnx = CUDART_INF, nmny = CUDART_INF, nmnz = CUDART_INF, nmxx = -CUDART_INF, nmxy = -CUDART_INF, nmxz = -CUDART_INF;
        double dx2 = 0.0;
        if (valid && !(pass == 1u && ghost)) {  // (ghosts sit out pass 1: see k_force_integrate)
            if (HIST && __popcll(amask_new) + __popc(wmask_new) > P.K)
                atomicOr(&C.err, ERR_HISTORY_OVERFLOW);
            const V3 x = integrate_store(P, B, C, src, dst, s, me, mv.v, mv.w, sid, flags, Fsum, Tsum, wmask_new, amask_new);
            if (!ghost) {
                nmnx = x.x - me.w; nmny = x.y - me.w; nmnz = x.z - me.w;
                nmxx = x.x + me.w; nmxy = x.y + me.w; nmxz = x.z + me.w;
            }
            const V3 dxv = x - mk(me.x, me.y, me.z);
            dx2 = dot(dxv, dxv);
        }
        {
            const bool grows = nmnx < dec_ord(C.bbox[0]) || nmny < dec_ord(C.bbox[1]) || nmnz < dec_ord(C.bbox[2]) ||
                               nmxx > dec_ord(C.bbox[3]) || nmxy > dec_ord(C.bbox[4]) || nmxz > dec_ord(C.bbox[5]);
            if (__any_sync(0xffffffffu, grows))
                block_bbox_commit(nmnx, nmny, nmnz, nmxx, nmxy, nmxz, C.bbox);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            dx2 = fmax(dx2, __shfl_xor_sync(0xffffffffu, dx2, o));
        if ((tid & 31) == 0) {
            const unsigned long long e = (unsigned long long)__double_as_longlong(dx2);
            if (e > C.max_dx2)
                atomicMax(&C.max_dx2, e);
        }
    }
    };  // rounds
    if (slow_block)
        rounds(std::true_type{});
    else
        rounds(std::false_type{});
}

// --------------------------------------------------------------------------------------------
// slab decomposition (one process per GPU).  All kernels below run only at a neighbour-list rebuild, except
// k_mgpu_pack / k_mgpu_unpack (the per-step halo) and k_mgpu_want.
// Record layouts (doubles): halo = pos xyz, v xyz, w xyz (9); ghost = pos4, VelRec (12);
// migrant = pos4, VelRec, acc[6], n_hist, K x (disp xyz, key|steps, relvel0)  (19 + 5 K).
// --------------------------------------------------------------------------------------------
constexpr int kHaloDoubles = 9;
constexpr int kGhostDoubles = 12;
__host__ __device__ inline int migrant_doubles(int K) { return 19 + 5 * K; }

// warp-aggregated slot allocation: one atomic per warp and counter
__device__ __forceinline__ unsigned warp_alloc(unsigned* counter, bool take) {
    const unsigned m = __ballot_sync(0xffffffffu, take);
    const unsigned lane = threadIdx.x & 31;
    unsigned base = 0;
    if (m && lane == (unsigned)(__ffs(m) - 1))
        base = atomicAdd(counter, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, m ? (__ffs(m) - 1) : 0);
    return base + __popc(m & ((1u << lane) - 1));
}

// live history of old slot `src` -> compact keyed records (same walk as k_gather_sorted)
template <class Emit>
__device__ __forceinline__ unsigned walk_history(const Params& P, const Buffers& B, const VelRec* vel_old, unsigned src,
                                                 unsigned meta, unsigned long long amask, Emit emit) {
    unsigned cnt = 0;
    unsigned wmask = (meta >> 8) & 0xFFFFu;
    while (wmask) {
        const int w = __ffs(wmask) - 1;
        wmask &= wmask - 1;
        const size_t si = (size_t)(P.Kn + w) * P.Np + src;
        if (cnt < (unsigned)P.K) {
            double4 r = B.hist[si];
            r.w = pack_key((unsigned)w, (unsigned)r.w);
            emit(cnt, r, B.hrel ? B.hrel[si] : 0.0);
        }
        cnt++;
    }
    while (amask) {
        const int k = __ffsll((long long)amask) - 1;
        amask &= amask - 1;
        const size_t si = (size_t)k * P.Np + src;
        if (cnt < (unsigned)P.K) {
            const unsigned jo = B.nl[si] & ~kEntryFlags;
            double4 r = B.hist[si];
            r.w = pack_key((jo & kTriFlag) ? (unsigned)P.nW + (jo & ~kTriFlag) : P.shape_base + vel_old[jo].sid, (unsigned)r.w);
            emit(cnt, r, B.hrel ? B.hrel[si] : 0.0);
        }
        cnt++;
    }
    return cnt;
}

// Step 1 of a slab rebuild: drop the ghosts, keep the spheres still inside [lo, hi) (compacted into the OTHER
// ping-pong buffer together with their staged history), write the ones that left into the migration messages.
__global__ void __launch_bounds__(256) k_mgpu_extract(Params P, Buffers B, double lo, double hi, double* out_left,
                                                      double* out_right, unsigned cap_out) {
    Ctrl& C = *B.ctrl;
    SlabDev& S = *B.slab;
    const unsigned a = C.cur, b = a ^ 1u;
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = s < P.N;
    double4 p = make_double4(0, 0, 0, 1);
    VelVal r;
    r.meta = 0; r.sid = 0; r.amask = 0; r.v = mk(0, 0, 0); r.w = mk(0, 0, 0);
    if (valid) {
        p = B.pos[a][s];
        r = load_vel(B.vel[a], s);
        if (r.meta & FLAG_GHOST)
            valid = false;
    }
    const int dest = !valid ? -1 : (p.x < lo ? 0 : (p.x >= hi ? 1 : 2));
    const unsigned ik = warp_alloc(&S.n_keep, dest == 2);
    const unsigned il = warp_alloc(&S.n_out[0], dest == 0);
    const unsigned ir = warp_alloc(&S.n_out[1], dest == 1);
    if (dest < 0)
        return;
    const unsigned flags = r.meta & 0xFFu;
    if (dest == 2) {
        B.pos[b][ik] = p;
        store_vel(B.vel[b], ik, r.v, r.w, r.sid, flags, 0ull);
        if (B.acc[0]) {
            const double2* as = reinterpret_cast<const double2*>(B.acc[a] + 6 * (size_t)s);
            double2* ad = reinterpret_cast<double2*>(B.acc[b] + 6 * (size_t)ik);
            ad[0] = as[0]; ad[1] = as[1]; ad[2] = as[2];
        }
        unsigned cnt = 0;
        if (B.hist) {
            cnt = walk_history(P, B, B.vel[a], s, r.meta, r.amask, [&](unsigned c, double4 h, double rel) {
                B.stage_init[(size_t)c * P.Np + ik] = h;
                if (B.stage_rel_init)
                    B.stage_rel_init[(size_t)c * P.Np + ik] = rel;
            });
            if (cnt > (unsigned)P.K) {
                atomicOr(&C.err, ERR_HISTORY_OVERFLOW);
                cnt = P.K;
            }
            B.stage_cnt_init[ik] = cnt;
        }
    } else {
        const unsigned io = dest == 0 ? il : ir;
        if (io >= cap_out) {
            atomicOr(&C.err, ERR_PAIR_CAPACITY);
            return;
        }
        if ((dest == 0 ? out_left : out_right) == nullptr) {  // left the outermost slab: there is nobody to take it
            atomicOr(&C.err, ERR_PAIR_CAPACITY);
            return;
        }
        double* o = (dest == 0 ? out_left : out_right) + (size_t)io * migrant_doubles(P.K);
        o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = p.w;
        o[4] = r.v.x; o[5] = r.v.y; o[6] = r.v.z; o[7] = r.w.x; o[8] = r.w.y; o[9] = r.w.z;
        o[10] = __hiloint2double((int)flags, (int)r.sid);
        o[11] = 0.0;
        for (int k = 0; k < 6; k++)
            o[12 + k] = B.acc[0] ? B.acc[a][6 * (size_t)s + k] : 0.0;
        unsigned cnt = 0;
        if (B.hist) {
            cnt = walk_history(P, B, B.vel[a], s, r.meta, r.amask, [&](unsigned c, double4 h, double rel) {
                double* q = o + 19 + 5 * c;
                q[0] = h.x; q[1] = h.y; q[2] = h.z; q[3] = h.w; q[4] = rel;
            });
            if (cnt > (unsigned)P.K)
                cnt = P.K;
        }
        o[18] = (double)cnt;
        __threadfence_system();  // the message buffer may be the neighbour's memory (P2P rebuild)
    }
}

// Append received records behind the kept spheres (same pre-sort arrays).  ghost != 0: light records, flagged.
__device__ __forceinline__ void append_record(const Params& P, const Buffers& B, const double* in, unsigned i, unsigned at,
                                              int ghost);

__global__ void __launch_bounds__(256) k_mgpu_append(Params P, Buffers B, const double* in, unsigned n, unsigned base,
                                                     int ghost) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    append_record(P, B, in, i, base + i, ghost);
}

__device__ __forceinline__ void append_record(const Params& P, const Buffers& B, const double* in, unsigned i, unsigned at,
                                              int ghost) {
    const Ctrl& C = *B.ctrl;
    const unsigned b = C.cur ^ 1u;
    const double* o = in + (size_t)i * (ghost ? kGhostDoubles : migrant_doubles(P.K));
    B.pos[b][at] = make_double4(o[0], o[1], o[2], o[3]);
    const unsigned sid = (unsigned)__double2loint(o[10]);
    unsigned flags = ((unsigned)__double2hiint(o[10])) & 0xFFu;
    flags = ghost ? (flags | FLAG_GHOST) : (flags & ~FLAG_GHOST);
    store_vel(B.vel[b], at, mk(o[4], o[5], o[6]), mk(o[7], o[8], o[9]), sid, flags, 0ull);
    if (ghost) {
        if (B.acc[0])
            for (int k = 0; k < 6; k++)
                B.acc[b][6 * (size_t)at + k] = 0.0;
        if (B.hist)
            B.stage_cnt_init[at] = 0;
        return;
    }
    if (B.acc[0])
        for (int k = 0; k < 6; k++)
            B.acc[b][6 * (size_t)at + k] = o[12 + k];
    if (B.hist) {
        const unsigned cnt = (unsigned)o[18];
        for (unsigned c = 0; c < cnt && c < (unsigned)P.K; c++) {
            const double* q = o + 19 + 5 * c;
            B.stage_init[(size_t)c * P.Np + at] = make_double4(q[0], q[1], q[2], q[3]);
            if (B.stage_rel_init)
                B.stage_rel_init[(size_t)c * P.Np + at] = q[4];
        }
        B.stage_cnt_init[at] = cnt;
    }
}

// Owned spheres (pre-sort indices [0, n_own)) within `cut` of a slab face: their copies go to that neighbour.
__global__ void __launch_bounds__(256) k_mgpu_select_ghosts(Params P, Buffers B, unsigned n_own, double lo, double hi,
                                                            double cut, double* out_left, double* out_right,
                                                            unsigned cap_out) {
    Ctrl& C = *B.ctrl;
    SlabDev& S = *B.slab;
    const unsigned b = C.cur ^ 1u;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_own == 0xFFFFFFFFu)
        n_own = S.n_own;  // device-driven rebuild: the count is only known on the device
    const bool valid = i < n_own;
    double4 p = make_double4(0, 0, 0, 1);
    if (valid)
        p = B.pos[b][i];
    const bool tl = valid && (p.x < lo + cut), tr = valid && (p.x >= hi - cut);
    const unsigned il = warp_alloc(&S.n_gsend[0], tl);
    const unsigned ir = warp_alloc(&S.n_gsend[1], tr);
    if (!tl && !tr)
        return;
    const VelVal r = load_vel(B.vel[b], i);
    for (int d = 0; d < 2; d++) {
        if (!(d == 0 ? tl : tr))
            continue;
        const unsigned io = d == 0 ? il : ir;
        if (io >= cap_out) {
            atomicOr(&C.err, ERR_PAIR_CAPACITY);
            continue;
        }
        B.send_pre[d][io] = i;
        double* o = (d == 0 ? out_left : out_right) + (size_t)io * kGhostDoubles;
        o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = p.w;
        o[4] = r.v.x; o[5] = r.v.y; o[6] = r.v.z; o[7] = r.w.x; o[8] = r.w.y; o[9] = r.w.z;
        o[10] = __hiloint2double((int)(r.meta & 0xFFu), (int)r.sid);
        o[11] = 0.0;
        __threadfence_system();  // the message buffer may be the neighbour's memory (P2P rebuild)
    }
}

// Last step of a slab rebuild: the assembled pre-sort arrays become the live buffer; the next step sorts them and
// rebuilds the candidate lists, taking the history from the staging columns.
__global__ void k_mgpu_finish(Buffers B) {
    if (threadIdx.x || blockIdx.x)
        return;
    Ctrl& C = *B.ctrl;
    C.cur ^= 1u;
    C.need_rebuild = 1u;
    C.init_stage = 1u;
    C.nrebuilds = 0ull;  // k_step_begin clears init_stage once nrebuilds >= 1: restart that latch
    C.travel = 0.0;
    C.travel_mesh = 0.0;
    C.max_dx2 = 0ull;
    SlabDev& S = *B.slab;
    S.n_keep = S.n_out[0] = S.n_out[1] = S.n_gsend[0] = S.n_gsend[1] = 0u;
    S.want_rebuild = 0u;
}

// After the sorting step: pre-sort indices -> storage slots for the per-step halo lists.
__global__ void __launch_bounds__(256) k_mgpu_invert_perm(Params P, Buffers B) {
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < P.N)
        B.inv_perm[B.perm[s]] = s;
}
__global__ void __launch_bounds__(256) k_mgpu_remap(Buffers B, unsigned n_own, unsigned ns0, unsigned ns1, unsigned ng0,
                                                    unsigned ng1) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    // the ghost senders are the second-pass set of the direct halo (kBndFlag): everything a neighbour needs of this slab is
    // computed behind the halo pick-up and can be sent while the first pass is still running
    if (i < ns0) {
        const unsigned sl = B.inv_perm[B.send_pre[0][i]];
        B.send_slot[0][i] = sl;
        atomicOr(&B.ncnt[sl], kBndFlag);
    }
    if (i < ns1) {
        const unsigned sl = B.inv_perm[B.send_pre[1][i]];
        B.send_slot[1][i] = sl;
        atomicOr(&B.ncnt[sl], kBndFlag);
    }
    if (i < ng0) B.ghost_slot[0][i] = B.inv_perm[n_own + i];
    if (i < ng1) B.ghost_slot[1][i] = B.inv_perm[n_own + ng0 + i];
}

// flagged spheres -> Buffers::bnd_list (each once, any order)
__global__ void __launch_bounds__(256) k_mgpu_bnd_list(Params P, Buffers B) {
    Ctrl& C = *B.ctrl;
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool take = s < P.N && (B.ncnt[s] & kBndFlag) != 0u;
    const unsigned at = warp_alloc(&C.n_bnd, take);
    if (take)
        B.bnd_list[at] = s;
}

// per-step halo: current state of the ghost-senders -> message; message -> ghost slots of the live buffer
__global__ void __launch_bounds__(256) k_mgpu_pack(Buffers B, int dir, unsigned n, double* out) {
    const Ctrl& C = *B.ctrl;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const unsigned s = B.send_slot[dir][i];
    const double4 p = B.pos[C.cur][s];
    const VelVal r = load_vel(B.vel[C.cur], s);
    double* o = out + (size_t)i * kHaloDoubles;
    o[0] = p.x; o[1] = p.y; o[2] = p.z;
    o[3] = r.v.x; o[4] = r.v.y; o[5] = r.v.z; o[6] = r.w.x; o[7] = r.w.y; o[8] = r.w.z;
}
__global__ void __launch_bounds__(256) k_mgpu_unpack(Buffers B, int dir, unsigned n, const double* in) {
    const Ctrl& C = *B.ctrl;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const unsigned s = B.ghost_slot[dir][i];
    const double* o = in + (size_t)i * kHaloDoubles;
    double4 p = B.pos[C.cur][s];
    p.x = o[0]; p.y = o[1]; p.z = o[2];
    B.pos[C.cur][s] = p;
    double2* q = reinterpret_cast<double2*>(B.vel[C.cur] + s);
    q[0] = make_double2(o[3], o[4]);
    q[1] = make_double2(o[5], o[6]);
    q[2] = make_double2(o[7], o[8]);
}
// Would the next step have to rebuild?  (travel so far + the displacement of the step that just ran)
// Travel since the last rebuild once the step that just ran and `ahead` more have been accounted for: those are assumed
// to move the spheres as far as the last one did, extrapolated linearly if the displacement per step is growing
// (accelerating bed), plus 25 %.
__device__ __forceinline__ double predicted_travel(const Ctrl& C, int ahead) {
    const double dx = sqrt(__longlong_as_double((long long)C.max_dx2));
    const double grow = fmax(dx - C.last_dx, 0.0);
    double t = C.travel + dx;
    for (int a = 1; a <= ahead; a++)
        t += 1.25 * (dx + (double)a * grow);
    return t;
}

// Same for the facet candidates: the mesh is assumed to keep moving as fast as its average since the last rebuild.
__device__ __forceinline__ bool mesh_travel_used_up(const Params& P, const Ctrl& C, int ahead) {
    if (!P.nT)
        return false;
    const double per_step = C.travel_mesh * 0.2;  // generous: a fifth of everything so far per further step
    return !(predicted_travel(C, ahead) + C.travel_mesh + (double)(1 + ahead) * per_step < 0.499 * P.skin_tri);
}

// `ahead` = further steps the caller will run before it acts on the answer (a driver that reads the flag one step late
// passes 1): they are assumed to move the spheres as far as the last step did; k_step_begin traps the case where that
// assumption fails (ERR_SKIN_EXCEEDED).
__global__ void k_mgpu_want(Params P, Buffers B, int* flag_out, int ahead) {
    if (threadIdx.x || blockIdx.x)
        return;
    const Ctrl& C = *B.ctrl;
    *flag_out = (C.need_rebuild != 0 || !(predicted_travel(C, ahead) < 0.499 * C.skin) || mesh_travel_used_up(P, C, ahead)) ? 1 : 0;
}
// --------------------------------------------------------------------------------------------
// direct P2P halo + vote (see P2PCtl in dem_types.h).  `step` = number of the time step the data belongs to = the
// value Ctrl::nsteps will have once k_step_begin of that step has run.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// dir 0: my spheres near my LEFT face -> left neighbour's "from the right" buffer; dir 1: the mirror image.
__global__ void __launch_bounds__(256) k_p2p_pack(Buffers B, P2PDev X, int dir, unsigned n) {
    const Ctrl& C = *B.ctrl;
    const unsigned long long step = C.nsteps + 1ull;
    double* out = X.peer_land[dir][step & 1ull];
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned s = B.send_slot[dir][i];
        const double4 p = B.pos[C.cur][s];
        const VelVal r = load_vel(B.vel[C.cur], s);
        double* o = out + (size_t)i * kHaloDoubles;
        o[0] = p.x; o[1] = p.y; o[2] = p.z;
        o[3] = r.v.x; o[4] = r.v.y; o[5] = r.v.z; o[6] = r.w.x; o[7] = r.w.y; o[8] = r.w.z;
    }
    // the last block to finish publishes the step number on the receiver.  One system fence per block: the barrier orders every
    // thread's stores before thread 0's fence, which is cumulative (PTX memory model: causality order through bar.sync)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned prev = atomicAdd(&X.done[dir], 1u);
        if (prev == gridDim.x - 1) {
            X.done[dir] = 0u;
            __threadfence_system();
            P2PCtl* peer = X.peer[X.rank + (dir == 0 ? -1 : 1)];
            // I am their right neighbour when I send to my left (dir 0), and vice versa
            st_release_sys(&peer->arrive[dir == 0 ? 1 : 0][step & 1ull], step);
        }
    }
}

// side 0: data from my left neighbour -> my ghost slots of side 0; side 1: from the right.
// after_begin = 0: runs in front of k_step_begin (the data belongs to the step about to begin, the live buffer is Ctrl::cur);
// 1: runs behind it, between the two passes of the force kernel (the step has begun: its number is Ctrl::nsteps, its input
// buffer Ctrl::f_src).
__global__ void __launch_bounds__(256) k_p2p_unpack(Buffers B, P2PDev X, int side, unsigned n, int after_begin) {
    const Ctrl& C = *B.ctrl;
    const unsigned long long step = C.nsteps + (after_begin ? 0ull : 1ull);
    const unsigned buf = after_begin ? C.f_src : C.cur;
    if (threadIdx.x == 0) {
        const unsigned long long* f = &X.self->arrive[side][step & 1ull];
        while (ld_acquire_sys(f) < step)
            __nanosleep(64);
    }
    __syncthreads();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const double* o = X.land[side][step & 1ull] + (size_t)i * kHaloDoubles;
    const unsigned s = B.ghost_slot[side][i];
    double4 p = B.pos[buf][s];
    p.x = __ldcv(o + 0); p.y = __ldcv(o + 1); p.z = __ldcv(o + 2);
    const double2 v01 = make_double2(__ldcv(o + 3), __ldcv(o + 4)), v2w0 = make_double2(__ldcv(o + 5), __ldcv(o + 6)),
                  w12 = make_double2(__ldcv(o + 7), __ldcv(o + 8));
    B.pos[buf][s] = p;
    double2* q = reinterpret_cast<double2*>(B.vel[buf] + s);
    q[0] = v01; q[1] = v2w0; q[2] = w12;
    if (after_begin) {
        // the step is under way and its first pass leaves the ghosts alone: their records of the output buffer are written here
        // (a ghost is never integrated: same state in both buffers; id, flags and radius come along)
        const double2 idm = q[3];
        B.pos[buf ^ 1u][s] = p;
        double2* qo = reinterpret_cast<double2*>(B.vel[buf ^ 1u] + s);
        qo[0] = v01; qo[1] = v2w0; qo[2] = w12; qo[3] = idm;
    }
}

// ---- device-driven rebuild (P2P mode): the same protocol as extract -> exchange -> append -> select_ghosts -> exchange
// -> append, but records are stored straight into the neighbour's landing buffers and the counts travel with them; the
// host only reads the final layout once.  `seq` = number of this rebuild.
__global__ void k_p2p_publish(Buffers B, P2PDev X, int what, unsigned long long seq) {
    if (threadIdx.x || blockIdx.x)
        return;
    const SlabDev& S = *B.slab;
    __threadfence_system();
    for (int d = 0; d < 2; d++) {
        const int nb = X.rank + (d == 0 ? -1 : 1);
        if (nb < 0 || nb >= X.world)
            continue;
        P2PCtl* peer = X.peer[nb];
        const int side = d == 0 ? 1 : 0;  // I am their right neighbour when I send to my left
        if (what == 0) {
            peer->mig_count[side] = S.n_out[d];
            __threadfence_system();
            st_release_sys(&peer->mig_arrive[side], seq);
        } else {
            peer->gho_count[side] = S.n_gsend[d];
            __threadfence_system();
            st_release_sys(&peer->gho_arrive[side], seq);
        }
    }
}

// what = 0: migrants of both sides behind the kept spheres; what = 1: ghosts behind the owned spheres (left first)
__global__ void __launch_bounds__(256) k_p2p_append(Params P, Buffers B, P2PDev X, int what, int side, unsigned long long seq) {
    SlabDev& S = *B.slab;
    Ctrl& C = *B.ctrl;
    const int nb = X.rank + (side == 0 ? -1 : 1);
    const bool have = nb >= 0 && nb < X.world;
    __shared__ unsigned s_n, s_base;
    if (threadIdx.x == 0) {
        unsigned n = 0;
        if (have) {
            const unsigned long long* f = what == 0 ? &X.self->mig_arrive[side] : &X.self->gho_arrive[side];
            while (ld_acquire_sys(f) < seq)
                __nanosleep(64);
            n = what == 0 ? X.self->mig_count[side] : X.self->gho_count[side];
        }
        unsigned base;
        if (what == 0)
            base = S.n_keep + (side == 1 ? S.n_in[0] : 0u);
        else
            base = S.n_own + (side == 1 ? S.g_in[0] : 0u);
        if ((what == 0 ? n > X.cap_mig : n > X.cap_gho) || (size_t)base + n > (size_t)P.Np) {
            atomicOr(&C.err, ERR_PAIR_CAPACITY);
            n = 0;
        }
        s_n = n;
        s_base = base;
    }
    __syncthreads();
    const unsigned n = s_n, base = s_base;
    const double* in = what == 0 ? X.mig_land[side] : X.gho_land[side];
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        append_record(P, B, in, i, base + i, what);
    // bookkeeping by the first thread of the grid, after which the next kernel in the stream reads it
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (what == 0) {
            S.n_in[side] = n;
            if (side == 1)
                S.n_own = S.n_keep + S.n_in[0] + n;
        } else {
            S.g_in[side] = n;
        }
    }
}

// After the step: my vote on "rebuild before the lists go stale?" goes to every rank; then the votes of the PREVIOUS
// step (every rank has long cast it) are OR-ed and handed to the host through mapped pinned memory.
__global__ void k_p2p_vote(Params P, Buffers B, P2PDev X) {
    const Ctrl& C = *B.ctrl;
    const unsigned long long step = C.nsteps;  // the step that just ran
    const unsigned r = threadIdx.x;
    if (r < (unsigned)X.world) {
        const unsigned long long flag =
            (C.need_rebuild != 0 || !(predicted_travel(C, X.ahead) < 0.499 * C.skin) || mesh_travel_used_up(P, C, X.ahead)) ? 1ull : 0ull;
        st_release_sys(&X.peer[r]->vote[step & 7ull][X.rank], (step << 1) | flag);
    }
    if (step <= X.first_step)
        return;
    const unsigned long long prev = step - 1ull;
    unsigned long long any = 0ull;
    if (r < (unsigned)X.world) {
        const unsigned long long* v = &X.self->vote[prev & 7ull][r];
        unsigned long long got;
        // a rank that has not run step `prev` in P2P mode yet (e.g. the very first step) never blocks us for long: the
        // host only acts on votes of steps it issued in P2P mode
        unsigned spins = 0;
        while (((got = ld_acquire_sys(v)) >> 1) < prev && ++spins < 4000000u)
            __nanosleep(64);
        any = ((got >> 1) == prev) ? (got & 1ull) : 1ull;  // timed out: ask for a rebuild rather than run on stale lists
    }
    any = __reduce_or_sync(0xffffffffu, (unsigned)any);
    if (r == 0) {
        X.host_vote[prev & 7ull] = (prev << 1) | any;
        __threadfence_system();
    }
}

// compact export of the owned spheres (any order): sid, pos, vel, omega
__global__ void __launch_bounds__(256) k_export_owned(Params P, Buffers B, unsigned* count, unsigned cap, unsigned* sid,
                                                      unsigned* slot_of, double* pos3, double* vel3, double* om3) {
    const Ctrl& C = *B.ctrl;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    bool own = false;
    double4 p = make_double4(0, 0, 0, 0);
    VelVal r;
    r.meta = 0; r.sid = 0; r.amask = 0; r.v = mk(0, 0, 0); r.w = mk(0, 0, 0);
    if (i < P.N) {
        p = B.pos[C.cur][i];
        r = load_vel(B.vel[C.cur], i);
        own = !(r.meta & FLAG_GHOST);
    }
    const unsigned at = warp_alloc(count, own);
    if (!own || at >= cap)
        return;
    sid[at] = r.sid;
    slot_of[at] = i;  // where record `at` of this export lives: dem_b200_import_owned hands the same records back
    if (pos3) { pos3[3 * (size_t)at] = p.x; pos3[3 * (size_t)at + 1] = p.y; pos3[3 * (size_t)at + 2] = p.z; }
    if (vel3) { vel3[3 * (size_t)at] = r.v.x; vel3[3 * (size_t)at + 1] = r.v.y; vel3[3 * (size_t)at + 2] = r.v.z; }
    if (om3) { om3[3 * (size_t)at] = r.w.x; om3[3 * (size_t)at + 1] = r.w.y; om3[3 * (size_t)at + 2] = r.w.z; }
}

// the inverse: records in the order of the last export -> their storage slots (no step in between: the order is unchanged)
__global__ void __launch_bounds__(256) k_import_owned(Buffers B, unsigned n, const unsigned* slot_of, const double* pos3,
                                                      const double* vel3, const double* om3) {
    Ctrl& C = *B.ctrl;
    const unsigned at = blockIdx.x * blockDim.x + threadIdx.x;
    if (at >= n)
        return;
    const unsigned i = slot_of[at];
    if (pos3) {
        double4 p = B.pos[C.cur][i];
        p.x = pos3[3 * (size_t)at]; p.y = pos3[3 * (size_t)at + 1]; p.z = pos3[3 * (size_t)at + 2];
        B.pos[C.cur][i] = p;
    }
    if (vel3 || om3) {
        VelVal r = load_vel(B.vel[C.cur], i);
        if (vel3) r.v = mk(vel3[3 * (size_t)at], vel3[3 * (size_t)at + 1], vel3[3 * (size_t)at + 2]);
        if (om3) r.w = mk(om3[3 * (size_t)at], om3[3 * (size_t)at + 1], om3[3 * (size_t)at + 2]);
        store_vel(B.vel[C.cur], i, r.v, r.w, r.sid, r.meta, r.amask);
    }
}

// ---- per-contact records (SetRecordingContactInfo): one pair, or all sphere-sphere contacts with bi < bj
// other_shape: shape id of the partner (wall w -> w, facet t -> nW + t, sphere j -> shape_base + j)
__global__ void __launch_bounds__(256) k_find_contact(Params P, Buffers B, unsigned sid_i, unsigned other_shape, double* out) {
    const Ctrl& C = *B.ctrl;
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.N)
        return;
    const VelRec* vel = B.vel[C.cur];
    if (vel[s].sid != sid_i || (vel[s].meta & FLAG_GHOST))
        return;
    int slot = -1;
    if (other_shape < (unsigned)P.nW) {
        if ((vel[s].meta >> (8 + other_shape)) & 1u)
            slot = P.Kn + (int)other_shape;
    } else {
        const unsigned nc = B.ncnt[s] & 0x7Fu, tc = B.ncnt[s] >> 24;
        for (unsigned k = 0; k < nc + tc; k++) {
            if (!((vel[s].amask >> k) & 1ull))
                continue;
            const unsigned e = B.nl[(size_t)k * P.Np + s] & ~kEntryFlags;
            const unsigned key = (e & kTriFlag) ? (unsigned)P.nW + (e & ~kTriFlag) : P.shape_base + vel[e].sid;
            if (key == other_shape)
                slot = (int)k;
        }
    }
    if (slot < 0)
        return;
    const double* src = B.cinfo + ((size_t)slot * P.Np + s) * kCInfo;
    for (int c = 0; c < kCInfo; c++)
        out[1 + c] = src[c];
    out[0] = 1.0;
}

__global__ void __launch_bounds__(256) k_export_contacts(Params P, Buffers B, unsigned* count, unsigned cap, unsigned* bi,
                                                         unsigned* bj, double* info) {
    const Ctrl& C = *B.ctrl;
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.N)
        return;
    const VelRec* vel = B.vel[C.cur];
    if (vel[s].meta & FLAG_GHOST)
        return;
    const unsigned nc = B.ncnt[s] & 0x7Fu;
    for (unsigned k = 0; k < nc; k++) {
        const unsigned e = B.nl[(size_t)k * P.Np + s];
        if (!(e & kHiFlag) || !((vel[s].amask >> k) & 1ull))
            continue;  // each pair once, from the lower id's side
        const unsigned at = atomicAdd(count, 1u);
        if (at >= cap)
            continue;
        bi[at] = vel[s].sid;
        bj[at] = vel[e & ~kEntryFlags].sid;
        const double* src = B.cinfo + ((size_t)k * P.Np + s) * kCInfo;
        for (int c = 0; c < kCInfo; c++)
            info[(size_t)at * kCInfo + c] = src[c];
    }
}

// --------------------------------------------------------------------------------------------
// reductions for the query API (GetMaxParticleZ, GetParticlesKineticEnergy, ...)
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_reduce(Params P, Buffers B, int which, double arg, double* out_sum,
                                                unsigned long long* out_ext) {
    const Ctrl& C = *B.ctrl;
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0, ext = -CUDART_INF;
    if (i < P.N) {
        double4 p = B.pos[C.cur][i];
        const double* vel = B.vel[C.cur][i].v;  // v[3] followed by w[3]
        switch (which) {
            case 0: ext = p.z; break;
            case 1: ext = -p.z; break;
            case 2: {
                double m = sphere_mass(P, p.w);
                double I = 0.4 * m * p.w * p.w;
                v = 0.5 * m * (vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]) +
                    0.5 * I * (vel[3] * vel[3] + vel[4] * vel[4] + vel[5] * vel[5]);
                break;
            }
            case 7: {  // translational part only: what Chrono::Dem's GetParticlesKineticEnergy sums (ChSystemDem_impl.cpp:1250-1264)
                double m = sphere_mass(P, p.w);
                v = 0.5 * m * (vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
                break;
            }
            case 3: ext = sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]); break;
            case 4: v = (p.z > arg) ? 1.0 : 0.0; break;
            case 5: v = (p.x > arg) ? 1.0 : 0.0; break;
            case 6:  // live contacts of the spheres this engine owns (a ghost's contacts are counted by its owner)
                v = (B.vel[C.cur][i].meta & FLAG_GHOST) ? 0.0
                    : (double)(__popcll(B.vel[C.cur][i].amask) + __popc((B.vel[C.cur][i].meta >> 8) & 0xFFFFu));
                break;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, o);
        ext = fmax(ext, __shfl_xor_sync(0xffffffffu, ext, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (which == 2 || which >= 4)
            atomicAdd(out_sum, v);
        else
            atomicMax(out_ext, enc_ord(ext));
    }
}

// user order <-> storage order
__global__ void __launch_bounds__(256) k_export_state(Params P, Buffers B, double* pos3, double* vel3, double* om3) {
    const Ctrl& C = *B.ctrl;
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N)
        return;
    const double4 p = B.pos[C.cur][i];
    const VelVal r = load_vel(B.vel[C.cur], i);
    const size_t o = 3 * (size_t)r.sid;
    if (pos3) { pos3[o] = p.x; pos3[o + 1] = p.y; pos3[o + 2] = p.z; }
    if (vel3) { vel3[o] = r.v.x; vel3[o + 1] = r.v.y; vel3[o + 2] = r.v.z; }
    if (om3) { om3[o] = r.w.x; om3[o + 1] = r.w.y; om3[o + 2] = r.w.z; }
}

// Linear acceleration of the last step, user order (GetParticleLinAcc, the fx,fy,fz columns of WriteParticleFile:
// ChSystemDem_impl.cpp:1290-1296, 322-327).  The fused force kernel does not keep the acceleration; after a step the
// other ping-pong buffer still holds the state the step started from, in the same slot order, and every integrator but
// Chung advances v by h * a, so a = (v+ - v) / h.  Chung keeps (a, alpha) of the step in B.acc.  Fixed spheres: 0.
__global__ void __launch_bounds__(256) k_export_accel(Params P, Buffers B, double* acc3) {
    const Ctrl& C = *B.ctrl;
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N)
        return;
    const VelVal r1 = load_vel(B.vel[C.cur], i);
    V3 a = mk(0.0, 0.0, 0.0);
    if (!(r1.meta & (FLAG_FIXED | FLAG_GHOST))) {
        if (P.integrator == 1) {
            const double* ap = B.acc[C.cur] + 6 * (size_t)i;
            a = mk(ap[0], ap[1], ap[2]);
        } else {
            const VelVal r0 = load_vel(B.vel[C.cur ^ 1u], i);
            const double inv_h = 1.0 / P.dt;
            a = mk((r1.v.x - r0.v.x) * inv_h, (r1.v.y - r0.v.y) * inv_h, (r1.v.z - r0.v.z) * inv_h);
        }
    }
    const size_t o = 3 * (size_t)r1.sid;
    acc3[o] = a.x; acc3[o + 1] = a.y; acc3[o + 2] = a.z;
}

__global__ void __launch_bounds__(256) k_import_state(Params P, Buffers B, const double* pos3, const double* vel3,
                                                      const double* om3) {
    Ctrl& C = *B.ctrl;
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N)
        return;
    VelVal r = load_vel(B.vel[C.cur], i);
    const size_t o = 3 * (size_t)r.sid;
    if (pos3) {
        double4 p = B.pos[C.cur][i];
        // how far the host moved this sphere: the candidate lists stay valid while the jumps fit into the Verlet skin
        // (Ctrl::import_dx2 -> k_step_begin); anything that is not a finite distance asks for the rebuild outright
        const double ex = pos3[o] - p.x, ey = pos3[o + 1] - p.y, ez = pos3[o + 2] - p.z;
        const double e2 = ex * ex + ey * ey + ez * ez;
        if (!(e2 < 1e300))
            C.need_rebuild = 1u;
        else if (e2 > 0.0) {
            const unsigned long long e = (unsigned long long)__double_as_longlong(e2);  // non-negative: raw bits are ordered
            if (e > *(volatile unsigned long long*)&C.import_dx2)  // (the running maximum soon filters out almost every thread)
                atomicMax(&C.import_dx2, e);
        }
        p.x = pos3[o]; p.y = pos3[o + 1]; p.z = pos3[o + 2];
        B.pos[C.cur][i] = p;
    }
    if (vel3) r.v = mk(vel3[o], vel3[o + 1], vel3[o + 2]);
    if (om3) r.w = mk(om3[o], om3[o + 1], om3[o + 2]);
    if (vel3 || om3)
        store_vel(B.vel[C.cur], i, r.v, r.w, r.sid, r.meta, r.amask);
}

}  // namespace demb200
