// =============================================================================
// dem_kernels.cuh -- hand-written sm_100a kernels of the SMC granular DEM step.
//
// One step (see dem_engine.cu for the launch order):
//   k_grid_update      1 thread    global AABB -> grid origin / bin size     (ChBroadphase.cpp:143-208)
//   k_bin_count        N threads   sphere -> bin of its AABB min corner, histogram (ChCollisionUtils.h:44-65)
//   k_scan_*           ncell       exclusive scan of the histogram (CSR bin starts)
//   k_scatter_perm     N threads   counting-sort permutation
//   k_gather_sorted    N threads   re-order posr/velw/sid by bin (deterministic order inside a bin: by sphere id)
//   k_force_integrate  N threads   narrowphase (sphere-sphere over 27 bins, sphere-wall), Hertz/Hooke/... force law
//                                  with pair-keyed tangential history, rolling/spinning resistance, gravity and
//                                  the time integrator -- fused, one pass over the state
// Everything that decides bin ids or contact-pair membership uses explicitly rounded fp64 intrinsics
// (__dmul_rn/__dadd_rn/__dsub_rn) so no FMA contraction can change a bit w.r.t. the Multicore arithmetic.
// =============================================================================
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
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

__global__ void __launch_bounds__(256) k_bbox_reduce(unsigned N, const double4* __restrict__ posr,
                                                     unsigned long long* bbox) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    double mnx = CUDART_INF, mny = CUDART_INF, mnz = CUDART_INF, mxx = -CUDART_INF, mxy = -CUDART_INF, mxz = -CUDART_INF;
    if (i < N) {
        double4 p = posr[i];
        mnx = p.x - p.w; mny = p.y - p.w; mnz = p.z - p.w;
        mxx = p.x + p.w; mxy = p.y + p.w; mxz = p.z + p.w;
    }
    block_bbox_commit(mnx, mny, mnz, mxx, mxy, mxz, bbox);
}

// --------------------------------------------------------------------------------------------
// grid of this step: ChBroadphase::DetermineBoundingBox + ComputeTopLevelResolution (ChBroadphase.cpp:143-208)
// --------------------------------------------------------------------------------------------
__global__ void k_grid_update(Params P, Buffers B) {
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    double mn[3], mx[3];
    for (int k = 0; k < 3; k++) {
        mn[k] = dec_ord(B.bbox[k]);
        mx[k] = dec_ord(B.bbox[3 + k]);
        if (P.has_wall_bb) {
            mn[k] = fmin(mn[k], P.wall_bb_min[k]);
            mx[k] = fmax(mx[k], P.wall_bb_max[k]);
        }
    }
    GridDev& G = *B.grid;
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
        if (!(bin >= 2.0 * P.rmax))
            err |= ERR_GRID_BIN_TOO_SMALL;
        if (!isfinite(lo) || !isfinite(hi))
            err |= ERR_NAN;
    }
    for (int w = 0; w < P.nW; w++)
        for (int k = 0; k < 3; k++) {
            G.wmin[w][k] = __dsub_rn(P.walls[w].amin[k], G.origin[k]);
            G.wmax[w][k] = __dsub_rn(P.walls[w].amax[k], G.origin[k]);
        }
    if (err)
        atomicOr(B.err, err);
    // restart the running sphere bounding box for the positions this step will produce
    for (int k = 0; k < 3; k++) {
        B.bbox[k] = P.has_wall_bb ? enc_ord(P.wall_bb_min[k]) : enc_ord(CUDART_INF);
        B.bbox[3 + k] = P.has_wall_bb ? enc_ord(P.wall_bb_max[k]) : enc_ord(-CUDART_INF);
    }
}

// HashMin of the sphere AABB lower corner, HashMax of the upper corner (ChCollisionUtils.h:44-60), computed on the
// origin-offset AABB exactly as OffsetAABB + f_Count_AABB_BIN_Intersection do.
struct BinRange {
    int lo[3], hi[3];
    double amin[3], amax[3];
};
__device__ __forceinline__ void sphere_bins(const double4& p, const double* org, const double* inv, BinRange& r) {
    const double c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        r.amin[k] = __dsub_rn(__dsub_rn(c[k], p.w), org[k]);
        r.amax[k] = __dsub_rn(__dadd_rn(c[k], p.w), org[k]);
        r.lo[k] = (int)floor(__dmul_rn(r.amin[k], inv[k]));
        r.hi[k] = (int)ceil(__dmul_rn(r.amax[k], inv[k])) - 1;
    }
}

// --------------------------------------------------------------------------------------------
// binning: bin id per sphere + histogram with warp-aggregated atomics
// --------------------------------------------------------------------------------------------
template <bool REC>
__global__ void __launch_bounds__(256) k_bin_count(Params P, Buffers B) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const GridDev& G = *B.grid;
    const bool valid = i < P.N;
    unsigned h = 0xFFFFFFFFu;
    if (valid) {
        double4 p = B.posA[i];
        BinRange r;
        sphere_bins(p, G.origin, G.inv, r);
        if (REC) {
            unsigned sid = B.sidA[i];
            for (int k = 0; k < 3; k++) {
                B.gmin[3 * sid + k] = r.lo[k];
                B.gmax[3 * sid + k] = r.hi[k];
            }
        }
        bool bad = false;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (r.lo[k] < 0 || r.lo[k] >= P.bins[k]) {
                bad = true;
                r.lo[k] = min(max(r.lo[k], 0), P.bins[k] - 1);
            }
        }
        if (bad)
            atomicOr(B.err, ERR_GRID_OUT_OF_RANGE);
        h = (unsigned)((r.lo[2] * P.bins[1] + r.lo[1]) * P.bins[0] + r.lo[0]);  // Hash_Index, z-major
        B.cell[i] = h;
    }
    // warp-aggregated histogram update: lanes that fall in the same bin elect a leader that issues one atomic;
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
// exclusive scan over the bin histogram: reduce-then-scan, 3 launches, tiles of 256 threads x 8 items
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

__global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(unsigned n, const uint32_t* __restrict__ in,
                                                                 uint32_t* __restrict__ tile_sums) {
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
        tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_sums(unsigned ntiles, uint32_t* tile_sums) {
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

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(unsigned n, unsigned total_items,
                                                             const uint32_t* __restrict__ in,
                                                             const uint32_t* __restrict__ tile_sums,
                                                             uint32_t* __restrict__ out) {
    const unsigned base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    unsigned v[kScanItems];
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    unsigned ex = block_exclusive_scan(s, nullptr) + tile_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n)
            out[base + k] = ex;
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        out[n] = total_items;
}

// --------------------------------------------------------------------------------------------
// counting sort: permutation, then gather of the 84-byte sphere records into bin order
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_scatter_perm(Params P, Buffers B) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N)
        return;
    B.perm[B.cell_start[B.cell[i]] + B.rank[i]] = i;
}

template <bool CHUNG>
__global__ void __launch_bounds__(256) k_gather_sorted(Params P, Buffers B) {
    const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= P.N)
        return;
    unsigned src = B.perm[s];
    const unsigned c = B.cell[src];
    const unsigned b = B.cell_start[c], e = B.cell_start[c + 1];
    if (e - b > 1) {
        // The atomics of k_bin_count leave an arbitrary order inside a bin.  Make it deterministic: slot b+r takes
        // the sphere with the r-th smallest stable id of the bin (bins hold a handful of spheres).
        const unsigned r = s - b;
        for (unsigned a = b; a < e; a++) {
            const unsigned ia = B.perm[a];
            const unsigned sa = B.sidA[ia];
            unsigned smaller = 0;
            for (unsigned j = b; j < e; j++)
                smaller += (B.sidA[B.perm[j]] < sa);
            if (smaller == r) {
                src = ia;
                break;
            }
        }
    }
    B.posB[s] = B.posA[src];
    const double2* vs = reinterpret_cast<const double2*>(B.velA + 6 * (size_t)src);
    double2* vd = reinterpret_cast<double2*>(B.velB + 6 * (size_t)s);
    vd[0] = vs[0];
    vd[1] = vs[1];
    vd[2] = vs[2];
    B.sidB[s] = B.sidA[src];
    if (CHUNG) {
        const double2* as = reinterpret_cast<const double2*>(B.accA + 6 * (size_t)src);
        double2* ad = reinterpret_cast<double2*>(B.accB + 6 * (size_t)s);
        ad[0] = as[0];
        ad[1] = as[1];
        ad[2] = as[2];
    }
}

// --------------------------------------------------------------------------------------------
// contact force law: ChIterativeSolverMulticoreSMC.cpp:56-546, bodies without orientation (spheres: the body
// frame can be taken parallel to the world frame), canonical orientation body1 = lower shape id.
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

template <bool HIST, bool ROLL>
__device__ __forceinline__ void contact_force(const Params& P, const Comp& cm, const Body& b1, const Body& b2,
                                              const Geom& g, Hist& h, V3& F, V3& T1, V3& T2) {
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

    if (ROLL) {
        double muRoll = cm.mu_roll, muSpin = cm.mu_spin;
        double d_coeff = gn_simple / (2.0 * m_eff * sqrt(kn_simple / m_eff));
        if (d_coeff < 1.0) {
            double t_collision = kPI * sqrt(m_eff / (kn_simple * (1 - d_coeff * d_coeff)));
            if (t_contact <= t_collision) {
                muRoll = 0.0;
                muSpin = 0.0;
            }
        }
        V3 v_rot = cross(b2.w, pt2_loc) - cross(b1.w, pt1_loc);
        V3 rel_o = b2.w - b1.w;
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
    }
    switch (P.adhesion_model) {
        case 0: force = force - cm.adh * g.n; break;
        case 1: force = force - cm.adh_dmt * sqrt(g.erad) * g.n; break;
        default: force = force - cm.adh_perko * g.erad * g.n; break;
    }
    F = force;
    T1 = tq1;
    T2 = tq2;
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

// --------------------------------------------------------------------------------------------
// fused narrowphase + force + integrate.  One thread per sphere in bin order.
// --------------------------------------------------------------------------------------------
constexpr int kForceThreads = 128;

template <bool HIST, bool ROLL, bool REC>
__global__ void __launch_bounds__(kForceThreads) k_force_integrate(const __grid_constant__ Params P,
                                                                   const __grid_constant__ Buffers B) {
    __shared__ unsigned clist[kMaxContactsPerSphere * kForceThreads];
    const unsigned tid = threadIdx.x;
    const unsigned s = blockIdx.x * kForceThreads + tid;
    const bool valid = s < P.N;
    const GridDev& G = *B.grid;

    double4 me = make_double4(0, 0, 0, 0);
    V3 mv = mk(0, 0, 0), mw = mk(0, 0, 0);
    unsigned sid = 0;
    int cnt = 0;
    BinRange br;
    if (valid) {
        me = B.posB[s];
        const double2* vp = reinterpret_cast<const double2*>(B.velB + 6 * (size_t)s);
        double2 a = vp[0], b = vp[1], c = vp[2];
        mv = mk(a.x, a.y, b.x);
        mw = mk(b.y, c.x, c.y);
        sid = B.sidB[s];
        sphere_bins(me, G.origin, G.inv, br);
        const int cx = min(max(br.lo[0], 0), P.bins[0] - 1);
        const int cy = min(max(br.lo[1], 0), P.bins[1] - 1);
        const int cz = min(max(br.lo[2], 0), P.bins[2] - 1);
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, P.bins[0] - 1);
        // ---- phase 1: candidate scan over the 3x3 rows of 3 contiguous bins; keep touching spheres ----
        for (int dz = -1; dz <= 1; dz++) {
            const int z = cz + dz;
            if (z < 0 || z >= P.bins[2])
                continue;
            for (int dy = -1; dy <= 1; dy++) {
                const int y = cy + dy;
                if (y < 0 || y >= P.bins[1])
                    continue;
                const unsigned row = (unsigned)((z * P.bins[1] + y) * P.bins[0]);
                const unsigned jb = B.cell_start[row + x0], je = B.cell_start[row + x1 + 1];
                for (unsigned j = jb; j < je; j++) {
                    if (j == s)
                        continue;
                    const double4 pj = B.posB[j];
                    // sphere_sphere test, ChNarrowphasePRIMS.cpp:50-59 (separation = 0)
                    const V3 d = mk(__dsub_rn(pj.x, me.x), __dsub_rn(pj.y, me.y), __dsub_rn(pj.z, me.z));
                    const double dist2 = dot_rn(d, d);
                    const double rs = __dadd_rn(me.w, pj.w);
                    if (dist2 >= __dmul_rn(rs, rs) || dist2 < 1e-12)
                        continue;
                    if (cnt < kMaxContactsPerSphere)
                        clist[cnt * kForceThreads + tid] = j;
                    cnt++;
                }
            }
        }
        if (cnt > kMaxContactsPerSphere) {
            atomicOr(B.err, ERR_CONTACT_LIST_OVERFLOW);
            cnt = kMaxContactsPerSphere;
        }
    }

    // ---- phase 2: all lanes evaluate their k-th contact together ----
    const double my_mass = sphere_mass(P, me.w);
    V3 Fsum = mk(0, 0, 0), Tsum = mk(0, 0, 0);
    int nh = 0;  // history slots written by this sphere
    unsigned ncontacts = 0;
    uint32_t* const my_hkey = HIST ? B.hkey_new + (size_t)sid * P.K : nullptr;
    double4* const my_hval = HIST ? B.hval_new + (size_t)sid * P.K : nullptr;
    const int maxc = __reduce_max_sync(0xffffffffu, cnt);
    for (int k = 0; k < maxc; k++) {
        if (k >= cnt)
            continue;
        const unsigned j = clist[k * kForceThreads + tid];
        const double4 pj = B.posB[j];
        const double2* vp = reinterpret_cast<const double2*>(B.velB + 6 * (size_t)j);
        const double2 a = vp[0], b = vp[1], c = vp[2];
        const unsigned sj = B.sidB[j];
        const bool me1 = sid < sj;  // canonical orientation: body 1 = lower shape id
        Body b1, b2;
        double r1, r2;
        {
            Body bm{mk(me.x, me.y, me.z), mv, mw, my_mass};
            Body bo{mk(pj.x, pj.y, pj.z), mk(a.x, a.y, b.x), mk(b.y, c.x, c.y), sphere_mass(P, pj.w)};
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
        ncontacts++;
        if (REC && !me1) {
            unsigned long long at = atomicAdd(B.pair_count, 1ull);
            if (at < B.pair_cap)
                B.pairs[at] = ((unsigned long long)(P.shape_base + sj) << 32) | (unsigned long long)(P.shape_base + sid);
        }
        if (g.depth >= 0)  // ChIterativeSolverMulticoreSMC.cpp:96-104: no force, no history
            continue;
        Hist h{mk(0, 0, 0), 0.0, 0.0, true};
        const unsigned key = P.shape_base + (me1 ? sid : sj);  // the non-owner's shape id
        if (HIST) {
            const size_t row = (size_t)(me1 ? sj : sid) * P.K;  // owner = higher id
            for (int t = 0; t < P.K; t++) {
                if (B.hkey_old[row + t] == key) {
                    const double4 hv = B.hval_old[row + t];
                    h.disp = mk(hv.x, hv.y, hv.z);
                    h.dur = hv.w;
                    if (B.hrel_old)
                        h.relvel0 = B.hrel_old[row + t];
                    h.isnew = false;
                    break;
                }
            }
        }
        V3 F, T1, T2;
        contact_force<HIST, ROLL>(P, P.comp[0], b1, b2, g, h, F, T1, T2);
        if (me1) {
            Fsum = Fsum - F;
            Tsum = Tsum + T1;
        } else {
            Fsum = Fsum + F;
            Tsum = Tsum + T2;
            if (HIST) {
                if (nh < P.K) {
                    my_hkey[nh] = key;
                    my_hval[nh] = make_double4(h.disp.x, h.disp.y, h.disp.z, h.dur);
                    if (B.hrel_new)
                        B.hrel_new[(size_t)sid * P.K + nh] = h.relvel0;
                }
                nh++;
            }
        }
    }

    double nmnx = CUDART_INF, nmny = CUDART_INF, nmnz = CUDART_INF, nmxx = -CUDART_INF, nmxy = -CUDART_INF, nmxz = -CUDART_INF;
    if (valid) {
        // ---- walls: body 1 = wall body (lower id), body 2 = this sphere, history on the sphere ----
        for (int w = 0; w < P.nW; w++) {
            const Wall& W = P.walls[w];
            Geom g;
            bool hit;
            if (W.type == WALL_BOX) {
                // broadphase AABB overlap on origin-offset boxes (ChCollisionUtils.h:83-87)
                if (!(br.amin[0] <= G.wmax[w][0] && G.wmin[w][0] <= br.amax[0] && br.amin[1] <= G.wmax[w][1] &&
                      G.wmin[w][1] <= br.amax[1] && br.amin[2] <= G.wmax[w][2] && G.wmin[w][2] <= br.amax[2]))
                    continue;
                hit = box_sphere_dev(W, mk(me.x, me.y, me.z), me.w, g);
            } else {
                hit = plane_sphere_dev(W, mk(me.x, me.y, me.z), me.w, g);
            }
            if (!hit)
                continue;
            ncontacts++;
            if (REC) {
                unsigned long long at = atomicAdd(B.pair_count, 1ull);
                if (at < B.pair_cap)
                    B.pairs[at] = ((unsigned long long)w << 32) | (unsigned long long)(P.shape_base + sid);
            }
            if (g.depth >= 0)
                continue;
            Hist h{mk(0, 0, 0), 0.0, 0.0, true};
            const unsigned key = (unsigned)w;
            if (HIST) {
                const size_t row = (size_t)sid * P.K;
                for (int t = 0; t < P.K; t++) {
                    if (B.hkey_old[row + t] == key) {
                        const double4 hv = B.hval_old[row + t];
                        h.disp = mk(hv.x, hv.y, hv.z);
                        h.dur = hv.w;
                        if (B.hrel_old)
                            h.relvel0 = B.hrel_old[row + t];
                        h.isnew = false;
                        break;
                    }
                }
            }
            Body b1{mk(0, 0, 0), mk(W.vel[0], W.vel[1], W.vel[2]), mk(0, 0, 0), P.wall_mass};
            Body b2{mk(me.x, me.y, me.z), mv, mw, my_mass};
            V3 F, T1, T2;
            contact_force<HIST, ROLL>(P, P.comp[1], b1, b2, g, h, F, T1, T2);
            Fsum = Fsum + F;
            Tsum = Tsum + T2;
            if (HIST) {
                if (nh < P.K) {
                    my_hkey[nh] = key;
                    my_hval[nh] = make_double4(h.disp.x, h.disp.y, h.disp.z, h.dur);
                    if (B.hrel_new)
                        B.hrel_new[(size_t)sid * P.K + nh] = h.relvel0;
                }
                nh++;
            }
        }
        if (HIST) {
            if (nh > P.K) {
                atomicOr(B.err, ERR_HISTORY_OVERFLOW);
                nh = P.K;
            }
            for (int t = nh; t < P.K; t++)
                my_hkey[t] = kEmptyKey;  // entries not touched this step are dropped (:677-685)
        }
        if (REC) {
            B.recF[3 * (size_t)sid + 0] = Fsum.x; B.recF[3 * (size_t)sid + 1] = Fsum.y; B.recF[3 * (size_t)sid + 2] = Fsum.z;
            B.recT[3 * (size_t)sid + 0] = Tsum.x; B.recT[3 * (size_t)sid + 1] = Tsum.y; B.recT[3 * (size_t)sid + 2] = Tsum.z;
            atomicAdd(B.n_contacts, (unsigned long long)ncontacts);
        }

        // ---- time integration ----
        const bool fixed = B.flags && (B.flags[sid] & 1);
        const double hdt = P.dt;
        const double inv_m = 1.0 / my_mass;
        const double inv_I = 1.0 / (0.4 * my_mass * me.w * me.w);
        const V3 gv = mk(P.g[0], P.g[1], P.g[2]);
        V3 x = mk(me.x, me.y, me.z);
        V3 vn = mv, wn = mw;
        if (!fixed) {
            if (P.integrator == 2) {
                // Multicore: hf = h*(g*m) + h*F; v+ = v + M^-1 hf; x+ = x + v+ h (ChBody.cpp:247-256,288-297)
                V3 hf = hdt * (gv * my_mass) + hdt * Fsum;
                vn = mv + inv_m * hf;
                wn = mw + inv_I * (hdt * Tsum);
                x = x + vn * hdt;
            } else {
                const V3 acc = gv + inv_m * Fsum;
                const V3 alp = inv_I * Tsum;
                if (P.integrator == 3) {         // extended Taylor (ChDemSMC.cuh:1347-1351)
                    x = x + hdt * (mv + 0.5 * hdt * acc);
                    vn = mv + hdt * acc;
                    wn = mw + hdt * alp;
                } else if (P.integrator == 0) {  // forward Euler
                    x = x + hdt * mv;
                    vn = mv + hdt * acc;
                    wn = mw + hdt * alp;
                } else {                         // Chung (ChDemSMC.cuh:1266-1277): beta = 28/27, gamma = 3/2
                    const double2* ap = reinterpret_cast<const double2*>(B.accB + 6 * (size_t)s);
                    const double2 o0 = ap[0], o1 = ap[1], o2 = ap[2];
                    const V3 ao = mk(o0.x, o0.y, o1.x), lo = mk(o1.y, o2.x, o2.y);
                    const double beta = 28.0 / 27.0;
                    x = x + hdt * (mv + hdt * (beta * acc + (0.5 - beta) * ao));
                    vn = mv + hdt * (1.5 * acc - 0.5 * ao);
                    wn = mw + hdt * (1.5 * alp - 0.5 * lo);
                    double2* aw = reinterpret_cast<double2*>(B.accA + 6 * (size_t)s);
                    aw[0] = make_double2(acc.x, acc.y);
                    aw[1] = make_double2(acc.z, alp.x);
                    aw[2] = make_double2(alp.y, alp.z);
                }
            }
        }
        if (!(isfinite(x.x) && isfinite(x.y) && isfinite(x.z)))
            atomicOr(B.err, ERR_NAN);
        B.posA[s] = make_double4(x.x, x.y, x.z, me.w);
        double2* vo = reinterpret_cast<double2*>(B.velA + 6 * (size_t)s);
        vo[0] = make_double2(vn.x, vn.y);
        vo[1] = make_double2(vn.z, wn.x);
        vo[2] = make_double2(wn.y, wn.z);
        B.sidA[s] = sid;
        nmnx = x.x - me.w; nmny = x.y - me.w; nmnz = x.z - me.w;
        nmxx = x.x + me.w; nmxy = x.y + me.w; nmxz = x.z + me.w;
    }
    // running bounding box of the new sphere AABBs -> next step's grid (all lanes take part in the shuffles)
    block_bbox_commit(nmnx, nmny, nmnz, nmxx, nmxy, nmxz, B.bbox);
}

// --------------------------------------------------------------------------------------------
// reductions for the query API (GetMaxParticleZ, GetParticlesKineticEnergy, ...)
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_reduce(Params P, Buffers B, int which, double arg, double* out_sum,
                                                unsigned long long* out_ext) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0, ext = -CUDART_INF;
    if (i < P.N) {
        double4 p = B.posA[i];
        const double* vel = B.velA + 6 * (size_t)i;
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
            case 3: ext = sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]); break;
            case 4: v = (p.z > arg) ? 1.0 : 0.0; break;
            case 5: v = (p.x > arg) ? 1.0 : 0.0; break;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, o);
        ext = fmax(ext, __shfl_xor_sync(0xffffffffu, ext, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (which == 2 || which == 4 || which == 5)
            atomicAdd(out_sum, v);
        else
            atomicMax(out_ext, enc_ord(ext));
    }
}

__global__ void __launch_bounds__(256) k_count_history(unsigned long long n, const uint32_t* __restrict__ keys,
                                                       unsigned long long* out) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned c = (i < n && keys[i] != kEmptyKey) ? 1u : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c)
        atomicAdd(out, (unsigned long long)c);
}

// user order <-> storage order
__global__ void __launch_bounds__(256) k_export_state(Params P, Buffers B, double* pos3, double* vel3, double* om3) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N)
        return;
    unsigned sid = B.sidA[i];
    double4 p = B.posA[i];
    const double* v = B.velA + 6 * (size_t)i;
    if (pos3) { pos3[3 * (size_t)sid] = p.x; pos3[3 * (size_t)sid + 1] = p.y; pos3[3 * (size_t)sid + 2] = p.z; }
    if (vel3) { vel3[3 * (size_t)sid] = v[0]; vel3[3 * (size_t)sid + 1] = v[1]; vel3[3 * (size_t)sid + 2] = v[2]; }
    if (om3) { om3[3 * (size_t)sid] = v[3]; om3[3 * (size_t)sid + 1] = v[4]; om3[3 * (size_t)sid + 2] = v[5]; }
}

__global__ void __launch_bounds__(256) k_import_state(Params P, Buffers B, const double* pos3, const double* vel3,
                                                      const double* om3) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N)
        return;
    unsigned sid = B.sidA[i];
    if (pos3) {
        double4 p = B.posA[i];
        p.x = pos3[3 * (size_t)sid]; p.y = pos3[3 * (size_t)sid + 1]; p.z = pos3[3 * (size_t)sid + 2];
        B.posA[i] = p;
    }
    double* v = B.velA + 6 * (size_t)i;
    if (vel3) { v[0] = vel3[3 * (size_t)sid]; v[1] = vel3[3 * (size_t)sid + 1]; v[2] = vel3[3 * (size_t)sid + 2]; }
    if (om3) { v[3] = om3[3 * (size_t)sid]; v[4] = om3[3 * (size_t)sid + 1]; v[5] = om3[3 * (size_t)sid + 2]; }
}

}  // namespace demb200
