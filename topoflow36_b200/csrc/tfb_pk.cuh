// Pieces shared by the packed routing kernels (tfb_channels_packed.cu: the lean
// production variants; tfb_channels_ext.cu: grid forcing, law of the wall, diagnostics):
// TMA + mbarrier plumbing, tile geometry, the peer-memory halo descriptor and the
// deterministic CTA / grid reduction.
#pragma once
#include "tfb_channels_common.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

namespace tfb {

constexpr unsigned kFull = 0xffffffffu;

enum { PK_SRC_NONE = 0, PK_SRC_GA = 1, PK_SRC_FULL = 2 };   // fused-source mode
enum { PK_KW = 0, PK_DWA = 1 };                              // main-kernel variant

struct PkConsts {
    double g, dt, da;
    double Runi;             // no fused source: the uniform runoff rate R
    int r_grid;              // accumulate vol_R per cell (pixel area varies by row)
    double Psum;             // ((P_rain + SM) + GW) + MR
    double P_total;          // P_rain + SM
    double Ks, gkq;          // Green-Ampt: Ks, (G*(Ks-Ki))*(qs-qi)
    int low_rain;            // P_total < Ks  (infil_base.py:954-995)
    int n_tiles, tiles_x;
    int tiles_y, n_boundary;     // tiles of the first and last tile row come FIRST in the order
    int reverse;                 // no peer halo: walk the tiles from the last one to the first
    // general fused sources (PK_SRC_FULL): every parameter a scalar, T_air may be a grid
    unsigned flags;
    int tair_grid;
    double P, SMv, GW, MR, ETv, INv, T_air, T_rs;
    double c0_864, T0, ratio;                    // c0 / 8.64e7, T0, rho_H2O / rho_snow
    double Qc, Qnet, alpha;                      // K_soil (T_surf - T_soil_x) / soil_x, Qn_SW + Qn_LW
};

constexpr int kRowGeom = 8;      // doubles per row: ds x3, 1/ds x3, da, pad


// ---------------------------------------------------------------------------
// TMA + mbarrier plumbing (sm_100a; SASS: UTMALDG / SYNCS)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned phase) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    return ok != 0u;
}
// bounded wait: a pipeline bug must not hang the GPU -- after ~2 s of polling the
// kernel records the failure and falls through (results are then flagged bad)
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, unsigned phase) {
    for (unsigned i = 0; i < (1u << 26); ++i) {
        if (mbar_try_wait(bar, phase)) return true;
    }
    return false;
}
// consumers of a peer-halo kernel: the producer may legitimately sit in peer_flag_wait for
// far longer than the bound above (a neighbour held up on its host); while it does
// (`*producer_waiting` != 0) the poll budget is not consumed
__device__ __forceinline__ bool mbar_wait_peer(uint64_t *bar, unsigned phase,
                                               const volatile int *producer_waiting,
                                               const volatile int *failed) {
    for (unsigned i = 0; i < (1u << 26); ++i) {
        if (mbar_try_wait(bar, phase)) return true;
        if (*failed) return false;               // the producer gave up: no tile will come
        if (*producer_waiting) i = 0u;
    }
    return false;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)),
           "r"(c0), "r"(c1) : "memory");
}

// tile geometry: 64 columns x R rows.  CPL = cells per lane: 2 -> one consumer warp
// per tile row (R + 1 warps, up to 128 registers per thread), 1 -> two warps per row
// (2R + 1 warps: more latency hiding, 64 registers).  The best choice depends on the
// register needs of the variant and is made per kernel instantiation (launch_main).
#ifndef TFB_PK_SMEM_KB
#define TFB_PK_SMEM_KB 196
#endif
constexpr int kSmemBudget = TFB_PK_SMEM_KB * 1024;   // dynamic shared memory for the ring
template <int ROWS, int CPL, int WIDTH = 64, int RPW = 1> struct TileCfg {
    static constexpr int R = ROWS * RPW;             // tile rows; a warp takes rows w, w + ROWS, ...
    static constexpr int RP = ROWS;                  // rows per pass of the consumer warps
    static constexpr int rpw = RPW;                  // passes: the per-tile bookkeeping (barrier
                                                     // wait, tile index, slot arithmetic) is paid
                                                     // once per RPW cells of a lane
    static constexpr int W = WIDTH;                  // tile columns
    static constexpr int QW = WIDTH + 4;             // halo box: cols c0-2 .. c0+W+1
    static constexpr int wpr = WIDTH / (32 * CPL);   // consumer warps per tile row
    static constexpr int QR = ROWS * RPW + 2;        // halo box: rows r0-1 .. r0+R
    static constexpr int cpl = CPL;
    static constexpr int warps = ROWS * wpr;         // consumer warps
    static constexpr int threads = (warps + 1) * 32; // + 1 producer warp
    static_assert(WIDTH == 64, "peer_push_tile moves 2 columns per lane of one warp");
    static constexpr int szQ = ((QW * QR * 8 + 127) / 128) * 128;
    static constexpr int szD = ((WIDTH * R * 8 + 127) / 128) * 128;   // one fp64 tile
    static constexpr int szF = ((WIDTH * R * 4 + 127) / 128) * 128;   // one 32-bit tile
    static constexpr int nD = WIDTH * R * 8, nF = WIDTH * R * 4, nQ = QW * QR * 8;  // TMA bytes
};

// v += x where the child-mask bit is set: a predicated DADD instead of add + select
__device__ __forceinline__ void add_if(double &v, unsigned m, unsigned bit, double x) {
    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\t"
        "@p add.rn.f64 %0, %0, %3;\n\t}" : "+d"(v) : "r"(m), "r"(bit), "d"(x));
}

// Processing order.  Row-major; with the peer-memory halo (PEER) the tiles that hold the
// slab's first and last owned rows come first (they are the only ones that read the halo
// rows and that feed the neighbours), then the rest.  The peer logic and this order cost
// 2.4 % on one GPU (same-box A/B), hence a template parameter and not a run-time flag.
template <bool PEER>
__host__ __device__ __forceinline__ void tile_of(const PkConsts &k, int s, int &ty, int &tx) {
    if constexpr (!PEER) {
        if (k.reverse) s = k.n_tiles - 1 - s;
        ty = s / k.tiles_x; tx = s - ty * k.tiles_x;
        return;
    }
    if (s < k.tiles_x) { ty = 0; tx = s; return; }
    if (s < k.n_boundary) { ty = k.tiles_y - 1; tx = s - k.tiles_x; return; }
    const int q = s - k.n_boundary;
    ty = 1 + q / k.tiles_x;
    tx = q - (ty - 1) * k.tiles_x;
}

// The same order walked WITHOUT a division per tile: position s advances by the grid size;
// the interior part (row-major from tile row 1, or from row 0 without the peer halo) is
// tracked incrementally, with one division when the walk enters it.
template <bool PEER> struct TileWalk {
    int ty, tx;                  // current tile
    int s, ity, itx;
    bool in;
    __host__ __device__ __forceinline__ void set(const PkConsts &k) {
        if constexpr (!PEER) { ty = ity; tx = itx; return; }
        if (s < k.tiles_x) { ty = 0; tx = s; return; }
        if (s < k.n_boundary) { ty = k.tiles_y - 1; tx = s - k.tiles_x; return; }
        if (!in) {
            const int q = s - k.n_boundary;
            ity = 1 + q / k.tiles_x;
            itx = q - (ity - 1) * k.tiles_x;
            in = true;
        }
        ty = ity; tx = itx;
    }
    __host__ __device__ __forceinline__ void init(const PkConsts &k, int s0) {
        s = s0; in = !PEER;
        const int i0 = (!PEER && k.reverse) ? (k.n_tiles - 1 - s0) : s0;    // may be < 0: no tile
        ity = PEER ? 0 : (i0 >= 0 ? i0 / k.tiles_x : 0);
        itx = PEER ? 0 : (i0 >= 0 ? i0 - ity * k.tiles_x : 0);
        set(k);
    }
    __host__ __device__ __forceinline__ void advance(const PkConsts &k, int step) {
        s += step;
        if (in) {
            if (!PEER && k.reverse) {
                itx -= step;
                while (itx < 0) { itx += k.tiles_x; --ity; }
            } else {
                itx += step;
                while (itx >= k.tiles_x) { itx -= k.tiles_x; ++ity; }
            }
        }
        set(k);
    }
};

constexpr uint32_t kPeerPoison = 0xffffffffu;   // flag value a failed rank leaves with its neighbours
struct PkPeer {
    double *up_Q, *down_Q;                  // the neighbours' halo rows of the new discharge
    uint32_t *up_flag, *down_flag;          // their "rows from rank r have arrived" flags
    uint32_t value;                         // published there when my boundary tiles are done
    const uint32_t *from_up, *from_down;    // my flags, written by the neighbours
    uint32_t need;                          // value they must hold before I touch a boundary tile
};

// Wait (bounded) until a neighbour has published step `need`: its halo row is in my memory
// and it no longer reads the halo row of its own that my boundary tiles overwrite.
__device__ __forceinline__ bool peer_flag_wait(const uint32_t *flag, uint32_t need) {
    // ~1 us per poll while spinning: gives a neighbour held up on its host (output, GC) about
    // a minute before the step is reported as failed instead of hanging the device
    for (unsigned i = 0; i < (1u << 26); ++i) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v == kPeerPoison) return false;          // the neighbour's step failed: fail fast
        if (v >= need) {
            // the halo row is read by the TMA unit (async proxy) next
            asm volatile("fence.proxy.async;" ::: "memory");
            return true;
        }
        __nanosleep(i < 1024u ? 64 : 512);
    }
    return false;
}

// CTA-level reduction of the per-thread accumulators, per-CTA row, last-CTA fold in
// CTA order (deterministic).  `mode`: 0 = everything, 1 = sums only (pass A), 2 =
// status only (pass B).
template <int NWARPS>
__device__ __forceinline__ void pk_finish(const tfb_channels_step_args &a, Acc &acc, float rate_max,
                                          bool consumer, const volatile int *failp, int mode,
                                          const PkPeer &peer,
                                          double (*sm)[kPartialStride], double *tot, bool *is_last) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (consumer) {
#pragma unroll
        for (int j = 0; j < kNSum; ++j) {
            double v = acc.s[j];
            if (__any_sync(kFull, v != 0.0)) v = warp_sum(v);
            if (lane == 0) sm[warp][j] = v;
        }
        unsigned any = 0;
#pragma unroll
        for (int j = 0; j < kNCnt; ++j) any |= acc.n[j];
        const bool some = __any_sync(kFull, any != 0u);
#pragma unroll
        for (int j = 0; j < kNCnt; ++j) {
            unsigned n = acc.n[j];
            if (some) n = __reduce_add_sync(kFull, n);
            if (lane == 0) sm[warp][kNSum + j] = (double)n;
        }
        const unsigned rb = __reduce_max_sync(kFull, __float_as_uint(fmaxf(rate_max, 0.0f)));
        if (lane == 0) sm[warp][P_DTMIN] = (double)__uint_as_float(rb);
    }
    __syncthreads();
    const int fail = *failp;                 // read AFTER the barrier: any warp may have set it
    if (tid < kPartialStride) {
        double v = 0.0;
        if (tid == P_DTMIN) {
            for (int w = 0; w < NWARPS; ++w) v = fmax(v, sm[w][tid]);
        } else if (tid < S_NSLOT) {
            for (int w = 0; w < NWARPS; ++w) v += sm[w][tid];
        }
        if (tid == P_FAIL) v = fail ? 1.0 : 0.0;          // this CTA gave up a bounded wait
        a.partials[(int64_t)(blockIdx.x + 2) * kPartialStride + tid] = v;
    }
    if (fail && tid == 0) {
        // the neighbours must not wait a minute each for rows that will never come
        if (peer.up_flag) *reinterpret_cast<volatile uint32_t *>(peer.up_flag) = kPeerPoison;
        if (peer.down_flag) *reinterpret_cast<volatile uint32_t *>(peer.down_flag) = kPeerPoison;
    }
    __threadfence();          // (peer halo rows were fenced system-wide by peer_push_tile)
    __syncthreads();
    if (tid == 0) {
        const unsigned tk = atomicAdd(a.ticket, 1u);
        *is_last = (tk == gridDim.x - 1u);
    }
    __syncthreads();
    if (!*is_last) return;
    __threadfence();
    if (tid < kPartialStride) {
        const volatile double *P = a.partials;
        double v = 0.0;
        if (tid < S_NSLOT || tid == P_DTMIN || tid == P_FAIL) {
            for (unsigned b = 0; b < gridDim.x; ++b) {
                const double x = P[(int64_t)(b + 2) * kPartialStride + tid];
                v = (tid == P_DTMIN) ? fmax(v, x) : (v + x);
            }
        }
        tot[tid] = v;
    }
    __syncthreads();
    if (tid == 0) {
        double *h = a.partials;                           // row 0: this rank's sums
        for (int j = 0; j < S_NSLOT; ++j) h[j] = tot[j];
        h[P_DTMIN] = (tot[P_DTMIN] > 0.0) ? (1.0 / tot[P_DTMIN]) : 3.0e38;
        double *outp = a.partials + kPartialStride;       // row 1: outlet staging
        const bool has_out = (outp[P_HASOUT] != 0.0);
        h[P_HASOUT_SUM] = has_out ? 1.0 : 0.0;
        h[P_QOUT] = has_out ? outp[P_QOUT] : 0.0;
        h[P_UOUT] = has_out ? outp[P_UOUT] : 0.0;
        h[P_DOUT] = has_out ? outp[P_DOUT] : 0.0;
        h[P_FOUT] = has_out ? outp[P_FOUT] : 0.0;
        h[P_HASOUT] = 0.0; h[P_FAIL] = tot[P_FAIL]; h[23] = 0.0;
        if (mode != 1) outp[P_HASOUT] = 0.0;
        if (a.finalize) {
            if (mode == 0) fold_scalars(a, h);
            else if (mode == 1) { fold_sums(a, h); fold_fail(a, h); }
            else { fold_status(a, h); fold_fail(a, h); mirror_scalars(a); }
        }
        a.ticket[0] = 0u;                                 // re-arm for the next launch
        a.ticket[1] = 0u;                                 // boundary-tile counter (peer halos)
    }
}

// TMA descriptor of a (rows, cols) array whose rows are `pitch` elements apart, box0 x box1
// elements per load; built on first use and cached (tfb_channels_packed.cu)
int pk_get_map(const void *ptr, int64_t rows, int64_t cols, int64_t pitch, CUtensorMapDataType dtype,
               int esize, int box0, int box1, CUtensorMap *out);

}  // namespace tfb
