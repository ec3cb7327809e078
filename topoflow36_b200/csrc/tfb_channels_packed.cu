// Packed fast path of the fused routing step (B200, sm_100a).
//
// Same physics as tfb_channels.cu -- channels_component.update()
// (components/channels_base.py:803-962) for the kinematic wave with Manning's
// formula (channels_kinematic_wave.py:84-127, channels_base.py:3960-4003),
// optionally with Green-Ampt infiltration fused in front of it
// (infil_green_ampt.py:511-578, infil_base.py:654-811, :954-995) -- specialised
// for the production configuration and organised around HBM bandwidth:
//
//   * static inputs come from the compact encoding of tfb_channels_packed
//     (16 B/cell instead of 42): geo32 word, float32 width, the Manning
//     coefficient sqrt|S|/n, and a <=256-entry bank-angle table in shared memory;
//   * one persistent CTA per SM: a producer warp streams 64 x 16-cell tiles of
//     every input (vol, Q with a 1-cell halo, coef, width, geo, I) into a 4-5
//     deep shared-memory ring with TMA (cp.async.bulk.tensor.2d + mbarrier
//     complete_tx), so ~130 KB of loads are in flight per SM at all times --
//     the Little's-law depth HBM3e needs -- independent of register pressure;
//     out-of-grid halo cells are zero-filled by the TMA unit, so there is no
//     edge code in the consumers;
//   * 16 consumer warps each take one row of the tile, two cells per lane: the
//     8 possible D8 children of a cell are read from the Q halo tile with
//     128-bit shared-memory loads and selected by the cell's mask bits -- the
//     inflow gather has no data-dependent global load; results leave with
//     128-bit coalesced global stores;
//   * tiles are dealt round-robin to the CTAs (16 Ki tiles at 4096^2: <0.3 %
//     imbalance), per-CTA partial sums are folded in CTA order by the last CTA:
//     the integrals are deterministic;
//   * divisions by static quantities are multiplications by correctly rounded
//     reciprocals; what is left per cell is one fp64 division (A_wet / P_wet),
//     one sqrt and one cbrt (+ one division for Green-Ampt).
//
// The children are added in the reference's order (codes 1,2,4,...,128 =
// child at SW, W, NW, N, NE, E, SE, S; channels_base.py:2116-2131).
#include "tfb_channels_common.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>
#include <vector>

namespace tfb {

constexpr unsigned kFull = 0xffffffffu;

struct PkConsts {
    double g, dt, da;
    double Rdadt;            // no fused source: (R*da)*dt, uniform
    double Psum;             // ((P_rain + SM) + GW) + MR
    double P_total;          // P_rain + SM
    double Ks, gkq;          // Green-Ampt: Ks, (G*(Ks-Ki))*(qs-qi)
    int low_rain;            // P_total < Ks  (infil_base.py:954-995)
    int n_tiles, tiles_x;
};

__global__ void pack_rows_kernel(const float *__restrict__ dx, const float *__restrict__ dy,
                                 const float *__restrict__ dd, double sinu, int ny,
                                 double *__restrict__ rg) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ny) return;
    const double ds0 = sinu * (double)dx[r], ds1 = sinu * (double)dd[r], ds2 = sinu * (double)dy[r];
    double *o = rg + (int64_t)r * 6;
    o[0] = ds0; o[1] = ds1; o[2] = ds2;
    o[3] = 1.0 / ds0; o[4] = 1.0 / ds1; o[5] = 1.0 / ds2;
}

__global__ void __launch_bounds__(256)
pack_geo32_kernel(const uint8_t *__restrict__ code, const uint8_t *__restrict__ aidx, int ny,
                  int nx, int halo, uint32_t *__restrict__ geo) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (c >= nx || r >= ny) return;
    const int64_t i = (int64_t)r * nx + c;
    // child position for mask bit k (= child's code 1<<k): opposite of direction k
    const int dr[8] = {1, 0, -1, -1, -1, 0, 1, 1};
    const int dc[8] = {-1, -1, -1, 0, 1, 1, 1, 0};
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int rr = r + dr[k], cc = c + dc[k];
        if (cc < 0 || cc >= nx || rr < -halo || rr >= ny + halo) continue;
        if (code[(int64_t)rr * nx + cc] == (uint8_t)(1u << k)) m |= (1u << k);
    }
    geo[i] = (uint32_t)code[i] | (m << 8) | ((uint32_t)(aidx ? aidx[i] : 0) << 16);
}

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
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)),
           "r"(c0), "r"(c1) : "memory");
}

// tile geometry: 64 columns x 16 rows; consumer warp w owns row w, a lane 2 cells
constexpr int kTW = 64, kTR = 16;
constexpr int kQW = kTW + 4, kQR = kTR + 2;          // dt*Q halo box (cols c0-2 .. c0+65)
constexpr int kConsumerWarps = kTR;
constexpr int kThreads = (kConsumerWarps + 1) * 32;  // + 1 producer warp
constexpr int kSzQ = ((kQW * kQR * 8 + 127) / 128) * 128;
constexpr int kSzD = kTW * kTR * 8;                  // one fp64 tile
constexpr int kSzF = kTW * kTR * 4;                  // one 32-bit tile

template <bool GA> struct StageLayout {
    static constexpr int oQ = 0;
    static constexpr int oVol = oQ + kSzQ;
    static constexpr int oCoef = oVol + kSzD;
    static constexpr int oI = oCoef + kSzD;
    static constexpr int oW = oI + (GA ? kSzD : 0);
    static constexpr int oGeo = oW + kSzF;
    static constexpr int bytes = oGeo + kSzF;
    static constexpr int tx_bytes = kQW * kQR * 8 + kSzD * (GA ? 3 : 2) + 2 * kSzF;
    static constexpr int stages = GA ? 4 : 5;
};

struct PkOut { double vol, d, u, Qn, I; };

template <bool GA>
__device__ __forceinline__ void pk_cell(const tfb_channels_step_args &a, const PkConsts &k,
                                        const double *__restrict__ lut, unsigned g, double vol,
                                        double Iv, double w, double coef, double qself,
                                        double c0_, double c1_, double c2_, double c3_,
                                        double c4_, double c5_, double c6_, double c7_,
                                        const double *__restrict__ rg, int r, int c,
                                        const tfb_channels_packed &p, Acc &acc, float &rate_max,
                                        PkOut &o) {
    const unsigned code = g & 0xffu, cmask = (g >> 8) & 0xffu, ai = (g >> 16) & 0xffu;
    const bool noflow = (code == 0);
    const int cls = (code & 0x55u) ? 1 : ((code & 0x88u) ? 2 : 0);
    const double ds = rg[cls], inv_ds = rg[3 + cls];
    const double dt = k.dt;
    double v;
    if (GA) {
        const double fc = (k.gkq / Iv) + k.Ks;                        // :540-542
        double IN = np_min(fc, k.P_total);                            // :567
        if (k.low_rain) IN = k.P_total;                               // infil_base :984-993
        o.I = Iv + (IN * dt);                                         // :797
        acc.s[S_VOL_IN] += (IN * k.da) * dt;                          // :732-736
        if (!isfinite(IN)) acc.n[S_NNAN_IN - kNSum] += 1u;
        const double R = k.Psum - (0.0 + IN);                         // channels_base :1649
        const double Rv = (R * k.da) * dt;
        acc.s[S_VOL_R] += Rv;                                         // :1666-1670
        v = vol + Rv;                                                 // :2086
    } else {
        v = vol + k.Rdadt;
    }
    if (cmask & 1u) v += dt * c0_;                                    // :2116-2131
    if (cmask & 2u) v += dt * c1_;
    if (cmask & 4u) v += dt * c2_;
    if (cmask & 8u) v += dt * c3_;
    if (cmask & 16u) v += dt * c4_;
    if (cmask & 32u) v += dt * c5_;
    if (cmask & 64u) v += dt * c6_;
    if (cmask & 128u) v += dt * c7_;
    v -= (qself * dt);                                                // :2138
    v = np_max(v, 0.0);                                               // :2145
    // update_channel_depth :2299-2330, :2379
    const double tn = lut[4 * ai], inv2tan = lut[4 * ai + 1], invcos = lut[4 * ai + 2];
    double d;
    if (tn == 0.0) {
        d = v / (w * ds);
    } else {
        double arg = (lut[4 * ai + 3] * v) * inv_ds;
        arg += w * w;
        d = (sqrt(arg) - w) * inv2tan;
    }
    d = np_max(d, 0.0);
    // update_trapezoid_Rh :2668-2703
    const double A_wet = d * (w + (d * tn));
    double P_wet = w + ((2.0 * d) * invcos);
    if (noflow) P_wet = 1.0;
    double Rh = A_wet / P_wet;
    if (noflow) Rh = 0.0;
    // manning_formula :3982-3983 with coef = sqrt|S_bed| / n
    const double cb = cbrt(Rh);
    double u = (cb * cb) * coef;
    if (noflow) u = 0.0;                                              // :2862
    if (r + a.row0 == a.outlet_row && c == a.outlet_col) {            // :2896-2904
        double f_o = 0.0;
        if (d > 0.0) f_o = k.g * ((p.outlet_rough * p.outlet_rough) / cbrt(d));
        double *po = a.partials + kPartialStride;
        po[P_QOUT] = qself; po[P_UOUT] = u; po[P_DOUT] = d; po[P_FOUT] = f_o; po[P_HASOUT] = 1.0;
    }
    o.Qn = u * A_wet;                                                 // next step :1697
    if (noflow) {                                                     // :2960-2966
        acc.s[S_VOL_EDGE] += v;
        v = 0.0;
        d = 0.0;
    }
    // check_flow_depth :3193 / check_flow_velocity :3352 -- rare path only
    if (!(d >= 0.0 && d <= 1.7976931348623157e308 && u >= 0.0 && u <= 1.7976931348623157e308)) {
        if (d < 0.0) acc.n[S_NNEG_D - kNSum] += 1u;
        if (d != d) acc.n[S_NNAN_D - kNSum] += 1u;
        if (isinf(d)) acc.n[S_NINF_D - kNSum] += 1u;
        if (u < 0.0) acc.n[S_NNEG_U - kNSum] += 1u;
        if (u != u) acc.n[S_NNAN_U - kNSum] += 1u;
        if (isinf(u)) acc.n[S_NINF_U - kNSum] += 1u;
    }
    rate_max = fmaxf(rate_max, (float)u * (float)inv_ds);             // CFL diagnostic: 1/dt_min
    o.vol = v; o.d = d; o.u = u;
}

struct PkMaps { CUtensorMap Q, vol, coef, I, w, geo; };

template <bool GA>
__global__ void __launch_bounds__(kThreads, 1)
channels_packed_kernel(const tfb_channels_step_args a, const tfb_channels_packed p,
                       const PkConsts k, const __grid_constant__ PkMaps maps) {
    using SL = StageLayout<GA>;
    constexpr int NS = SL::stages;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ double lut[256 * 4];
    __shared__ double sm[kConsumerWarps][kPartialStride];
    __shared__ __align__(8) uint64_t full_bar[NS], empty_bar[NS];
    __shared__ bool is_last;
    __shared__ int s_fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < p.n_angles * 4; i += kThreads) lut[i] = __ldg(p.angle_lut + i);
    if (tid < kConsumerWarps * kPartialStride) (&sm[0][0])[tid] = 0.0;
    if (tid == 0) {
        s_fail = 0;
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full_bar[s], 1u);
            mbar_init(&empty_bar[s], (unsigned)kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    Acc acc;
#pragma unroll
    for (int j = 0; j < kNSum; ++j) acc.s[j] = 0.0;
#pragma unroll
    for (int j = 0; j < kNCnt; ++j) acc.n[j] = 0u;
    float rate_max = 0.0f;

    if (warp == kConsumerWarps) {
        // ---------------- producer: one lane streams tiles into the ring -------------
        if (lane == 0) {
            int stage = 0;
            unsigned phase = 0;
            for (int t = blockIdx.x; t < k.n_tiles; t += gridDim.x) {
                if (!mbar_wait(&empty_bar[stage], phase ^ 1u)) { s_fail = 1; break; }
                const int ty = t / k.tiles_x, tx = t - ty * k.tiles_x;
                const int c0 = tx * kTW, r0 = ty * kTR;
                unsigned char *st = smem + (size_t)stage * SL::bytes;
                mbar_arrive_expect_tx(&full_bar[stage], (unsigned)SL::tx_bytes);
                // tensor rows are counted from the top of the halo'd allocation
                tma_load_2d(st + SL::oQ, &maps.Q, &full_bar[stage], c0 - 2, r0 - 1 + a.halo);
                tma_load_2d(st + SL::oVol, &maps.vol, &full_bar[stage], c0, r0 + a.halo);
                tma_load_2d(st + SL::oCoef, &maps.coef, &full_bar[stage], c0, r0);
                if (GA) tma_load_2d(st + SL::oI, &maps.I, &full_bar[stage], c0, r0 + a.halo);
                tma_load_2d(st + SL::oW, &maps.w, &full_bar[stage], c0, r0);
                tma_load_2d(st + SL::oGeo, &maps.geo, &full_bar[stage], c0, r0);
                if (++stage == NS) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // ---------------- consumers: warp w computes row w of every tile -------------
        int stage = 0;
        unsigned phase = 0;
        for (int t = blockIdx.x; t < k.n_tiles; t += gridDim.x) {
            if (!mbar_wait(&full_bar[stage], phase)) { s_fail = 1; break; }
            const int ty = t / k.tiles_x, tx = t - ty * k.tiles_x;
            const int c0 = tx * kTW + lane * 2, r = ty * kTR + warp;
            const unsigned char *st = smem + (size_t)stage * SL::bytes;
            if (r < a.ny && c0 < a.nx) {
                const int e = warp * kTW + lane * 2;                   // element in a 64x16 tile
                const uint2 geo = *reinterpret_cast<const uint2 *>(st + SL::oGeo + e * 4);
                const float2 w = *reinterpret_cast<const float2 *>(st + SL::oW + e * 4);
                const double2 vol = *reinterpret_cast<const double2 *>(st + SL::oVol + e * 8);
                const double2 coef = *reinterpret_cast<const double2 *>(st + SL::oCoef + e * 8);
                double2 Iv = make_double2(0.0, 0.0);
                if (GA) Iv = *reinterpret_cast<const double2 *>(st + SL::oI + e * 8);
                // Q window: halo rows warp, warp+1, warp+2; halo cols 2*lane+1 .. 2*lane+4
                const double *q = reinterpret_cast<const double *>(st + SL::oQ) + warp * kQW + lane * 2;
                const double2 a0 = *reinterpret_cast<const double2 *>(q);
                const double2 a1 = *reinterpret_cast<const double2 *>(q + 2);
                const double2 m0 = *reinterpret_cast<const double2 *>(q + kQW);
                const double2 m1 = *reinterpret_cast<const double2 *>(q + kQW + 2);
                const double2 b0 = *reinterpret_cast<const double2 *>(q + 2 * kQW);
                const double2 b1 = *reinterpret_cast<const double2 *>(q + 2 * kQW + 2);
                // columns: x0 = c0-2 (unused), a0.y = c0-1, a1.x = c0, a1.y = c0+1; the
                // east neighbour c0+2 of cell 1 is the first value of the next pair
                const double aE = q[4], mE = q[kQW + 4], bE = q[2 * kQW + 4];
                const double *rg = p.row_geom + (int64_t)r * 6;
                PkOut o0, o1;
                o0.I = o1.I = 0.0;
                // cell 0 (col c0): west = *.0.y, same = *.1.x, east = *.1.y
                pk_cell<GA>(a, k, lut, geo.x, vol.x, Iv.x, (double)w.x, coef.x, m1.x,
                            b0.y, m0.y, a0.y, a1.x, a1.y, m1.y, b1.y, b1.x,
                            rg, r, c0, p, acc, rate_max, o0);
                // cell 1 (col c0+1): west = *.1.x, same = *.1.y, east = *E
                pk_cell<GA>(a, k, lut, geo.y, vol.y, Iv.y, (double)w.y, coef.y, m1.y,
                            b1.x, m1.x, a1.x, a1.y, aE, mE, bE, b1.y,
                            rg, r, c0 + 1, p, acc, rate_max, o1);
                const int64_t i = (int64_t)r * a.nx + c0;
                *reinterpret_cast<double2 *>(a.vol_out + i) = make_double2(o0.vol, o1.vol);
                *reinterpret_cast<double2 *>(a.d + i) = make_double2(o0.d, o1.d);
                *reinterpret_cast<double2 *>(a.u + i) = make_double2(o0.u, o1.u);
                *reinterpret_cast<double2 *>(a.Q_out + i) = make_double2(o0.Qn, o1.Qn);
                if (GA) *reinterpret_cast<double2 *>(a.I_out + i) = make_double2(o0.I, o1.I);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);            // slot may be refilled
            if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
        // ---- per-CTA partial sums: warp shuffles only where a warp has something -------
        double e = acc.s[S_VOL_EDGE];
        if (__any_sync(kFull, e != 0.0)) e = warp_sum(e);
        if (lane == 0) sm[warp][S_VOL_EDGE] = e;
        if (GA) {
            const double vr = warp_sum(acc.s[S_VOL_R]), vi = warp_sum(acc.s[S_VOL_IN]);
            if (lane == 0) { sm[warp][S_VOL_R] = vr; sm[warp][S_VOL_IN] = vi; }
        }
        unsigned any = 0;
#pragma unroll
        for (int j = 0; j < kNCnt; ++j) any |= acc.n[j];
        if (__any_sync(kFull, any != 0u)) {
#pragma unroll
            for (int j = 0; j < kNCnt; ++j) {
                const unsigned n = __reduce_add_sync(kFull, acc.n[j]);
                if (lane == 0) sm[warp][kNSum + j] = (double)n;
            }
        }
        const unsigned rb = __reduce_max_sync(kFull, __float_as_uint(fmaxf(rate_max, 0.0f)));
        if (lane == 0) sm[warp][P_DTMIN] = (double)__uint_as_float(rb);
    }
    __syncthreads();
    if (tid < kPartialStride) {
        double v = 0.0;
        if (tid == P_DTMIN) {
            for (int w = 0; w < kConsumerWarps; ++w) v = fmax(v, sm[w][tid]);
        } else if (tid < S_NSLOT) {
            for (int w = 0; w < kConsumerWarps; ++w) v += sm[w][tid];
        }
        if (tid == S_NNAN_D && s_fail) v += 1.0e9;       // pipeline failure poisons the step
        a.partials[(int64_t)(blockIdx.x + 2) * kPartialStride + tid] = v;
    }
    // ---- completion: the last CTA folds the per-CTA rows in CTA order ------------------
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned tk = atomicAdd(a.ticket, 1u);
        is_last = (tk == gridDim.x - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    __shared__ double tot[kPartialStride];
    if (tid < kPartialStride) {
        const volatile double *P = a.partials;
        double v = 0.0;
        if (tid < S_NSLOT || tid == P_DTMIN) {
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
        h[P_HASOUT] = 0.0; h[22] = 0.0; h[23] = 0.0;
        outp[P_HASOUT] = 0.0;
        if (a.finalize) fold_scalars(a, h);
        a.ticket[0] = 0u;                                 // re-arm for the next step
    }
}

}  // namespace tfb

using namespace tfb;

extern "C" int64_t tfb_sizeof_channels_packed(void) { return (int64_t)sizeof(tfb_channels_packed); }

extern "C" int tfb_channels_pack_geo32(const uint8_t *code, const uint8_t *angle_idx, int32_t ny,
                                       int32_t nx, int32_t halo, uint32_t *geo32, void *stream) {
    TFB_ARG_CHECK(code && geo32 && ny > 0 && nx > 0 && halo >= 0);
    dim3 b(32, 8), g((nx + 31) / 32, (ny + 7) / 8);
    pack_geo32_kernel<<<g, b, 0, (cudaStream_t)stream>>>(code, angle_idx, ny, nx, halo, geo32);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_channels_pack_rows(const float *dx, const float *dy, const float *dd,
                                      double sinu, int32_t ny, double *row_geom, void *stream) {
    TFB_ARG_CHECK(dx && dy && dd && row_geom && ny > 0);
    pack_rows_kernel<<<(ny + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dx, dy, dd, sinu, ny,
                                                                       row_geom);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

// ---- TMA descriptors: built on first use, cached by (address, shape, box) --------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct MapKey {
    const void *ptr; int64_t rows, cols; int dtype; int box0, box1; int dev;
    bool operator==(const MapKey &o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && dtype == o.dtype &&
               box0 == o.box0 && box1 == o.box1 && dev == o.dev;
    }
};
struct MapEntry { MapKey key; CUtensorMap map; };
std::mutex g_map_mu;
std::vector<MapEntry> g_maps;
EncodeTiledFn g_encode = nullptr;

int get_map(const void *ptr, int64_t rows, int64_t cols, CUtensorMapDataType dtype, int esize,
            int box0, int box1, CUtensorMap *out) {
    int dev = 0;
    TFB_CUDA_CHECK(cudaGetDevice(&dev));
    const MapKey key{ptr, rows, cols, (int)dtype, box0, box1, dev};
    std::lock_guard<std::mutex> lock(g_map_mu);
    for (const MapEntry &e : g_maps)
        if (e.key == key) { *out = e.map; return TFB_OK; }
    if (!g_encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        TFB_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault,
                                               &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) {
            tfb::set_error("cuTensorMapEncodeTiled not available from the driver");
            return TFB_ERR_CUDA;
        }
        g_encode = (EncodeTiledFn)fn;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)cols * (cuuint64_t)esize};
    const cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    const cuuint32_t estr[2] = {1u, 1u};
    MapEntry e;
    e.key = key;
    const CUresult rc = g_encode(&e.map, dtype, 2u, const_cast<void *>(ptr), gdim, gstride, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) {
        tfb::set_error("cuTensorMapEncodeTiled failed (%d) for %p rows %lld cols %lld", (int)rc, ptr,
                       (long long)rows, (long long)cols);
        return TFB_ERR_CUDA;
    }
    if (g_maps.size() > 256) g_maps.clear();
    g_maps.push_back(e);
    *out = e.map;
    return TFB_OK;
}

template <bool GA>
int launch_packed(const tfb_channels_step_args &a, const tfb_channels_packed &p, const PkConsts &k,
                  cudaStream_t st) {
    using SL = StageLayout<GA>;
    PkMaps m;
    const int64_t rows_h = (int64_t)a.ny + 2 * a.halo;
    const size_t hb = (size_t)a.halo * a.nx;               // elements between allocation and row 0
    int rc;
    if ((rc = get_map(a.Q_in - hb, rows_h, a.nx, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, kQW, kQR, &m.Q))) return rc;
    if ((rc = get_map(a.vol - hb, rows_h, a.nx, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, kTW, kTR, &m.vol))) return rc;
    if ((rc = get_map(p.coef, a.ny, a.nx, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, kTW, kTR, &m.coef))) return rc;
    if (GA) {
        if ((rc = get_map(a.I - hb, rows_h, a.nx, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, kTW, kTR, &m.I))) return rc;
    } else {
        m.I = m.vol;
    }
    if ((rc = get_map(p.width32, a.ny, a.nx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, kTW, kTR, &m.w))) return rc;
    if ((rc = get_map(p.geo32, a.ny, a.nx, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, kTW, kTR, &m.geo))) return rc;
    const int smem = SL::stages * SL::bytes;
    static bool attr_set[2][64] = {{false}};
    int dev = 0;
    TFB_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[GA ? 1 : 0][dev]) {
        TFB_CUDA_CHECK(cudaFuncSetAttribute(channels_packed_kernel<GA>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (dev < 64) attr_set[GA ? 1 : 0][dev] = true;
    }
    int64_t grid = sm_count();
    if (grid > k.n_tiles) grid = k.n_tiles;
    channels_packed_kernel<GA><<<(int)grid, kThreads, smem, st>>>(a, p, k, m);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

}  // namespace

extern "C" int tfb_channels_step_packed(const tfb_channels_step_args *ap,
                                        const tfb_channels_packed *pp, void *stream) {
    TFB_ARG_CHECK(ap != nullptr && pp != nullptr);
    const tfb_channels_step_args &a = *ap;
    const tfb_channels_packed &p = *pp;
    TFB_ARG_CHECK(a.ny > 0 && a.nx > 0 && a.halo >= 0 && (a.nx % 4) == 0);
    TFB_ARG_CHECK(p.geo32 && p.width32 && p.coef && p.angle_lut && p.row_geom && p.n_angles >= 1 &&
                  p.n_angles <= 256);
    TFB_ARG_CHECK(a.dx && a.dy && a.dd && a.vol && a.vol_out && a.Q_in && a.Q_out && a.d && a.u);
    TFB_ARG_CHECK(a.Q_in != a.Q_out && a.scalars && a.partials && a.ticket && a.dt > 0.0);
    // eligibility: kinematic + Manning, lean outputs, uniform forcing
    const uint32_t src = a.flags & (TFB_CH_SRC_MET | TFB_CH_SRC_SNOW | TFB_CH_SRC_EVAP |
                                    TFB_CH_SRC_INFIL);
    TFB_ARG_CHECK(!(a.flags & (TFB_CH_DIFFUSIVE | TFB_CH_LAW_OF_WALL | TFB_CH_DIAG |
                               TFB_CH_WRITE_R)));
    TFB_ARG_CHECK(src == 0 || src == TFB_CH_SRC_INFIL);
    TFB_ARG_CHECK(!a.sinu.ptr && !a.da.ptr);
    TFB_ARG_CHECK(!a.P_rain.ptr && !a.SM.ptr && !a.GW.ptr && !a.MR.ptr && !a.ET.ptr);
    auto al16 = [](const void *q) { return (((uintptr_t)q) & 15u) == 0; };
    TFB_ARG_CHECK(al16(a.vol) && al16(a.vol_out) && al16(a.Q_in) && al16(a.Q_out) && al16(a.d) &&
                  al16(a.u) && al16(p.coef) && al16(p.row_geom) && ((((uintptr_t)p.geo32) & 7u) == 0) &&
                  ((((uintptr_t)p.width32) & 7u) == 0));
    PkConsts k;
    k.g = 9.81;
    k.dt = a.dt;
    k.da = a.da.val;
    k.Psum = ((a.P_rain.val + a.SM.val) + a.GW.val) + a.MR.val;
    k.P_total = a.P_rain.val + a.SM.val;
    k.Ks = 0.0; k.gkq = 0.0; k.low_rain = 0;
    k.tiles_x = (a.nx + kTW - 1) / kTW;
    k.n_tiles = k.tiles_x * ((a.ny + kTR - 1) / kTR);
    cudaStream_t st = (cudaStream_t)stream;
    if (src) {
        TFB_ARG_CHECK(a.I && a.I_out && al16(a.I) && al16(a.I_out) && a.is_R_grid == 1);
        TFB_ARG_CHECK(!a.Ks.ptr && !a.Ki.ptr && !a.qs.ptr && !a.qi.ptr && !a.G.ptr &&
                      !a.h_table.ptr && !a.elev.ptr);
        TFB_ARG_CHECK(!(a.h_table.val >= a.elev.val));          // infil_base.py:697-701
        k.Ks = a.Ks.val;
        k.gkq = (a.G.val * (a.Ks.val - a.Ki.val)) * (a.qs.val - a.qi.val);
        k.low_rain = (k.P_total < k.Ks) ? 1 : 0;
        k.Rdadt = 0.0;
        return launch_packed<true>(a, p, k, st);
    } else {
        TFB_ARG_CHECK(!a.IN.ptr && a.is_R_grid == 0);
        const double R = k.Psum - (0.0 + a.IN.val);
        k.Rdadt = (R * a.da.val) * a.dt;
        return launch_packed<false>(a, p, k, st);
    }
}
