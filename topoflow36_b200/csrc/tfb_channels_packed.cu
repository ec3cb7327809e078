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
#include "tfb_pk.cuh"

#include <string.h>

#include <mutex>
#include <vector>

namespace tfb {

__global__ void pack_rows_kernel(const float *__restrict__ dx, const float *__restrict__ dy,
                                 const float *__restrict__ dd, double sinu,
                                 const double *__restrict__ da_row, double da, int ny,
                                 double *__restrict__ rg) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= ny) return;
    const double ds0 = sinu * (double)dx[r], ds1 = sinu * (double)dd[r], ds2 = sinu * (double)dy[r];
    double *o = rg + (int64_t)r * kRowGeom;
    o[0] = ds0; o[1] = ds1; o[2] = ds2;
    o[3] = 1.0 / ds0; o[4] = 1.0 / ds1; o[5] = 1.0 / ds2;
    o[6] = da_row ? da_row[r] : da;
    o[7] = 0.0;
}

__global__ void __launch_bounds__(256)
pack_geo32_kernel(const uint8_t *__restrict__ code, const uint16_t *__restrict__ aidx, int ny,
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

// TFB_PK_RG_TMA: the per-row geometry (flow lengths, reciprocals, pixel area: 8 doubles per
// row) rides the ring as one more TMA box instead of being fetched by the consumers with
// global loads at the head of every tile (the kernel's only long-scoreboard stalls)
#ifndef TFB_PK_RG_TMA
#define TFB_PK_RG_TMA 1      // same-box A/B: KW + Green-Ampt 0.254 -> 0.236 ms, KW 0.224 -> 0.216
#endif
// ring slot of the main kernel (kinematic step, or pass A of the diffusive step)
template <int SRC, int VAR, class TC> struct StageLayout {
    static constexpr int kSzQ = TC::szQ, kSzD = TC::szD, kSzF = TC::szF;
    static constexpr bool kCoef = (VAR == PK_KW);
    static constexpr bool kI = (SRC != PK_SRC_NONE);
    static constexpr bool kFullSrc = (SRC == PK_SRC_FULL);
    static constexpr int oQ = 0;
    static constexpr int oVol = oQ + kSzQ;
    static constexpr int oCoef = oVol + kSzD;
    static constexpr int oI = oCoef + (kCoef ? kSzD : 0);
    static constexpr int oH = oI + (kI ? kSzD : 0);          // h_swe   (FULL only)
    static constexpr int oT = oH + (kFullSrc ? kSzD : 0);    // T_air   (FULL only)
    static constexpr int oW = oT + (kFullSrc ? kSzD : 0);
    static constexpr int oGeo = oW + kSzF;
    static constexpr int oRG = oGeo + kSzF;
    static constexpr int nRG = TC::R * kRowGeom * 8;                       // TMA bytes of the row box
    static constexpr int bytes = oRG + (TFB_PK_RG_TMA ? ((nRG + 127) / 128) * 128 : 0);
    static constexpr int stages = (kSmemBudget / bytes) > 6 ? 6 : (kSmemBudget / bytes);
};
// ring slot of pass B of the diffusive step
template <class TC> struct StageLayoutB {
    static constexpr int kSzQ = TC::szQ, kSzD = TC::szD, kSzF = TC::szF;
    static constexpr int oD = 0;                             // new depth with halo
    static constexpr int oN = oD + kSzQ;                     // 1 / n
    static constexpr int oS = oN + kSzD;                     // S_bed (f32)
    static constexpr int oW = oS + kSzF;
    static constexpr int oGeo = oW + kSzF;
    static constexpr int oRG = oGeo + kSzF;
    static constexpr int nRG = TC::R * kRowGeom * 8;
    static constexpr int bytes = oRG + (TFB_PK_RG_TMA ? ((nRG + 127) / 128) * 128 : 0);
    static constexpr int tx_bytes = TC::nQ + TC::nD + 3 * TC::nF + (TFB_PK_RG_TMA ? nRG : 0);
    static constexpr int stages = (kSmemBudget / bytes) > 6 ? 6 : (kSmemBudget / bytes);
};

struct PkIn { unsigned g; double vol, I, h, Ta, w, qself, ch[8]; };

struct PkVD { double v, d, I, h; };    // new volume / depth before the noflow zeroing, new state

// new volume and depth of one cell: update_R :1548, update_flow_volume :1951,
// update_channel_depth :2250 (+ the fused source terms in front of them).  STRAIGHT-LINE
// code (selects, predicated adds, the Newton primitives of tfb_common.cuh): the calls for
// the two cells a lane advances land in ONE basic block, so ptxas interleaves their
// dependent fp64 chains.
template <int SRC>
__device__ __forceinline__ void pk_vol_depth(const PkConsts &k, const double *__restrict__ lut,
                                             const PkIn &x, const double ds, const double inv_ds,
                                             const double da, Acc &acc, PkVD &o) {
    const unsigned cmask = (x.g >> 8) & 0xffu, ai = (x.g >> 16) & 0xffu;
    const double dt = k.dt;
    double v;
    o.I = 0.0; o.h = 0.0;
    if (SRC == PK_SRC_GA) {
        const double fc = div_sel(k.gkq, x.I) + k.Ks;                 // infil_green_ampt :540-542
        double IN = np_min(fc, k.P_total);                            // :567
        IN = k.low_rain ? k.P_total : IN;                             // infil_base :984-993
        o.I = x.I + (IN * dt);                                        // :797
        acc.s[S_VOL_IN] += (IN * da) * dt;                            // :732-736
        acc.n[S_NNAN_IN - kNSum] += isfinite(IN) ? 0u : 1u;
        const double R = k.Psum - (0.0 + IN);                         // channels_base :1649
        const double Rv = (R * da) * dt;
        acc.s[S_VOL_R] += Rv;                                         // :1666-1670
        v = x.vol + Rv;                                               // :2086
    } else if (SRC == PK_SRC_FULL) {
        // every source term is evaluated and the disabled ones are dropped by selects on the
        // (kernel-uniform) flags: no branch, and configs[3] enables all four anyway
        const bool f_met = (k.flags & TFB_CH_SRC_MET) != 0u, f_snow = (k.flags & TFB_CH_SRC_SNOW) != 0u;
        const bool f_evap = (k.flags & TFB_CH_SRC_EVAP) != 0u, f_inf = (k.flags & TFB_CH_SRC_INFIL) != 0u;
        const double Ta = k.tair_grid ? x.Ta : k.T_air;
        // met_base :1300-1388
        const double Pr_m = k.P * ((Ta > k.T_rs) ? 1.0 : 0.0);
        const double Ps_m = k.P * ((Ta <= k.T_rs) ? 1.0 : 0.0);
        const double P_rain = f_met ? Pr_m : k.P;
        const double P_snow = f_met ? Ps_m : 0.0;
        acc.s[S_VOL_P] += f_met ? ((k.P * da) * dt) : 0.0;
        acc.s[S_VOL_PR] += f_met ? ((P_rain * da) * dt) : 0.0;
        acc.s[S_VOL_PS] += f_met ? ((P_snow * da) * dt) : 0.0;
        // snow_degree_day :293-360, snow_base :490-512, :559-609
        double M = k.c0_864 * (Ta - k.T0);
        M = np_min(M, div_nr(x.h, dt));
        M = np_max0(M);
        const double SM = f_snow ? M : k.SMv;
        double hn = x.h + (P_snow * dt);
        hn = hn - (M * dt);
        o.h = np_max0(hn);
        acc.s[S_VOL_SM] += f_snow ? ((M * da) * dt) : 0.0;
        // evap_priestley_taylor :308-413
        const double Qet = (k.alpha * (0.406 + (0.011 * Ta))) * (k.Qnet + k.Qc);
        double ET = np_max0(div_nr(Qet, 2.5E+9));
        acc.s[S_VOL_ET] += f_evap ? ((ET * da) * dt) : 0.0;
        // infil_green_ampt :511-578
        const double P_total = P_rain + SM;
        const double fc = div_sel(k.gkq, x.I) + k.Ks;
        double INg = np_min(fc, P_total);
        INg = (P_total < k.Ks) ? P_total : INg;
        const double IN = f_inf ? INg : k.INv;
        o.I = x.I + (INg * dt);
        acc.s[S_VOL_IN] += f_inf ? ((INg * da) * dt) : 0.0;
        acc.n[S_NNAN_IN - kNSum] += (f_inf && !isfinite(INg)) ? 1u : 0u;
        // channels_base :1625-1629: a grid ET is masked where the cell cannot supply it, a
        // scalar ET (no evaporation component fused) is dropped
        ET = (f_evap && !(x.vol < ((ET * da) * dt))) ? ET : 0.0;
        const double R = (((P_rain + SM) + k.GW) + k.MR) - (ET + IN);
        const double Rv = (R * da) * dt;
        acc.s[S_VOL_R] += Rv;
        v = x.vol + Rv;
    } else {
        const double Rv = (k.Runi * da) * dt;
        acc.s[S_VOL_R] += k.r_grid ? Rv : 0.0;                        // :1666-1670 (da varies)
        v = x.vol + Rv;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) add_if(v, cmask, 1u << j, dt * x.ch[j]);   // :2116-2131
    v -= (x.qself * dt);                                              // :2138
    v = np_max0(v);                                                   // :2145
    // update_channel_depth :2299-2330, :2379; table entry = {tan, 1/(2 tan), 1/cos, 4 tan}.
    // Both forms are evaluated: the trapezoid one, and v / (w ds) for vertical banks
    const double2 t01 = *reinterpret_cast<const double2 *>(lut + 4 * ai);
    double arg = (lut[4 * ai + 3] * v) * inv_ds;
    arg += x.w * x.w;
    const double d_trap = (sqrt_sel(arg) - x.w) * t01.y;
    const double d_rect = div_sel(v, x.w * ds);
    const double d = (t01.x == 0.0) ? d_rect : d_trap;
    o.v = v;
    o.d = np_max0(d);
}

// hydraulic radius, velocity, next discharge of one cell (straight-line): update_trapezoid_Rh
// :2643, manning_formula :3960, update_velocity_on_edges :2851.  Returns "a stability check
// trips on this cell" (check_flow_depth :3193, check_flow_velocity :3352): the caller counts.
__device__ __forceinline__ bool pk_velocity(const double *__restrict__ lut, unsigned g, double d,
                                            double w, double umul, double inv_ds,
                                            float &rate_max, double &u_out, double &Qn_out) {
    const unsigned code = g & 0xffu, ai = (g >> 16) & 0xffu;
    const bool noflow = (code == 0);
    const double tn = lut[4 * ai], invcos = lut[4 * ai + 2];          // {tan, 1/(2 tan), 1/cos, 4 tan}
    const double A_wet = d * (w + (d * tn));                          // :2668-2703
    double P_wet = w + ((2.0 * d) * invcos);
    P_wet = noflow ? 1.0 : P_wet;
    double Rh = div_nr(A_wet, P_wet);
    Rh = noflow ? 0.0 : Rh;
    double u = pow23_sel(Rh) * umul;                                  // :3982-3983
    u = noflow ? 0.0 : u;                                             // :2862
    Qn_out = u * A_wet;                                               // next step :1697
    const double dz = noflow ? 0.0 : d;                               // :2960-2966
    rate_max = fmaxf(rate_max, (float)u * (float)inv_ds);             // CFL diagnostic: 1/dt_min
    u_out = u;
    return !(dz >= 0.0 && dz <= 1.7976931348623157e308 && u >= 0.0 && u <= 1.7976931348623157e308);
}
// the rare tails, outside the straight-line block --------------------------------------------
__device__ __forceinline__ void pk_count_bad(Acc &acc, unsigned g, double d, double u) {
    const double dz = ((g & 0xffu) == 0u) ? 0.0 : d;
    if (dz < 0.0) acc.n[S_NNEG_D - kNSum] += 1u;
    if (dz != dz) acc.n[S_NNAN_D - kNSum] += 1u;
    if (isinf(dz)) acc.n[S_NINF_D - kNSum] += 1u;
    if (u < 0.0) acc.n[S_NNEG_U - kNSum] += 1u;
    if (u != u) acc.n[S_NNAN_U - kNSum] += 1u;
    if (isinf(u)) acc.n[S_NINF_U - kNSum] += 1u;
}
__device__ __forceinline__ void pk_outlet(const tfb_channels_step_args &a, double g, double qself,
                                       double u, double d, double outlet_rough) {   // :2896-2904
    double f_o = 0.0;
    if (d > 0.0) f_o = g * ((outlet_rough * outlet_rough) / cbrt(d));
    double *po = a.partials + kPartialStride;
    po[P_QOUT] = qself; po[P_UOUT] = u; po[P_DOUT] = d; po[P_FOUT] = f_o; po[P_HASOUT] = 1.0;
}

__global__ void zero_at_kernel(double *__restrict__ x, const int64_t *__restrict__ idx, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[__ldg(idx + i)] = 0.0;
}

struct PkMaps { CUtensorMap Q, vol, coef, I, H, T, w, geo, rg; };
struct PkMapsB { CUtensorMap D, N, S, w, geo, rg; };

// Peer-memory halo, sending side -- executed by the PRODUCER warp (all 32 lanes), off the
// consumers' critical path: when the ring slot of boundary tile `t` has been released
// (every consumer warp has stored its results and arrived on the slot's `empty` barrier:
// release / acquire at CTA scope), the warp copies the tile's share of the slab's first /
// last owned row of `src` (the new discharge, or the new pre-zero depth in pass A of the
// diffusive step) from local memory into the neighbour's halo row over NVLink, makes the
// stores visible system-wide and counts the tile; the warp that completes the count
// publishes the step in both neighbours' flag words.
template <class TC>
__device__ __forceinline__ void peer_push_tile(const tfb_channels_step_args &a, const PkConsts &k,
                                               const PkPeer &peer, const double *__restrict__ src,
                                               int ty, int txi, int lane) {
    const int c = txi * TC::W + 2 * lane;                  // TC::W = 64 columns: 2 per lane
    if (c < a.nx) {
        const bool two = (c + 1 < a.nx);
        if (ty == 0 && peer.up_Q) {
            const double *sp = src + c;
            if (two) *reinterpret_cast<double2 *>(peer.up_Q + c) = __ldcg(reinterpret_cast<const double2 *>(sp));
            else peer.up_Q[c] = __ldcg(sp);
        }
        if (ty == k.tiles_y - 1 && peer.down_Q) {
            const double *sp = src + (int64_t)(a.ny - 1) * a.pitch + c;
            if (two) *reinterpret_cast<double2 *>(peer.down_Q + c) = __ldcg(reinterpret_cast<const double2 *>(sp));
            else peer.down_Q[c] = __ldcg(sp);
        }
    }
    __threadfence_system();                                // my lanes' peer stores are visible
    __syncwarp();
    if (lane == 0 && atomicAdd(a.ticket + 1, 1u) == (unsigned)k.n_boundary - 1u) {
        __threadfence_system();                            // ... and so are the other CTAs' (counted)
        if (peer.up_flag) *reinterpret_cast<volatile uint32_t *>(peer.up_flag) = peer.value;
        if (peer.down_flag) *reinterpret_cast<volatile uint32_t *>(peer.down_flag) = peer.value;
    }
}

// ---------------------------------------------------------------------------
// main kernel: the kinematic step (VAR = PK_KW) or pass A of the diffusive step
// (VAR = PK_DWA: new volume, new depth and source state only)
// ---------------------------------------------------------------------------
template <int SRC, int VAR, class TC, bool PEER>
__global__ void __launch_bounds__(TC::threads, 1)
channels_packed_kernel(const tfb_channels_step_args a, const tfb_channels_packed p,
                       const PkConsts k, const __grid_constant__ PkMaps maps, const PkPeer peer) {
    using SL = StageLayout<SRC, VAR, TC>;
    constexpr int kTR = TC::R, kTW = TC::W, kQW = TC::QW, kConsumerWarps = TC::warps, kThreads = TC::threads;
    constexpr int NS = SL::stages;
    constexpr bool kKW = (VAR == PK_KW);
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(16) double lut[256 * 4];
    __shared__ double sm[kConsumerWarps][kPartialStride];
    __shared__ double tot[kPartialStride];
    __shared__ __align__(8) uint64_t full_bar[NS], empty_bar[NS];
    __shared__ bool is_last;
    __shared__ volatile int s_fail;
    __shared__ volatile int s_pwait;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < p.n_angles * 4; i += kThreads) lut[i] = __ldg(p.angle_lut + i);
    if (tid < kConsumerWarps * kPartialStride) (&sm[0][0])[tid] = 0.0;
    if (tid == 0) {
        s_fail = 0;
        s_pwait = 0;
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
    const bool use_h = (SRC == PK_SRC_FULL) && (k.flags & TFB_CH_SRC_SNOW);
    const bool use_i = (SRC == PK_SRC_GA) || ((SRC == PK_SRC_FULL) && (k.flags & TFB_CH_SRC_INFIL));
    const bool use_t = (SRC == PK_SRC_FULL) && k.tair_grid;

    if (warp == kConsumerWarps) {
        // ---------------- producer warp: lane 0 streams tiles into the ring; with the
        // peer-memory halo the whole warp also forwards finished boundary rows -------------
        int stage = 0;
        unsigned phase = 0;
        bool up_ok = false, down_ok = false;
        const unsigned tx = (unsigned)(TC::nQ + TC::nD * (1 + (kKW ? 1 : 0) + (use_i ? 1 : 0) +
                                                          (use_h ? 1 : 0) + (use_t ? 1 : 0)) + 2 * TC::nF +
                                       (TFB_PK_RG_TMA ? SL::nRG : 0));
        const double *push_src = kKW ? a.Q_out : a.d;
        // PEER: NS extra (load-free) turns drain the boundary tiles still in the ring
        const int t_end = k.n_tiles + (PEER ? NS * (int)gridDim.x : 0);
        for (int t = blockIdx.x; t < t_end; t += gridDim.x) {
            const int t_prev = t - NS * (int)gridDim.x;           // tile that held this slot before
            const bool push = PEER && t_prev >= 0 && t_prev < k.n_boundary;
            const bool load = t < k.n_tiles;
            if (load || push) {
                int ok = 1;
                if (lane == 0) ok = mbar_wait(&empty_bar[stage], phase ^ 1u) ? 1 : 0;
                ok = __shfl_sync(kFull, ok, 0);
                __syncwarp();                 // lane 0's acquire orders the other lanes' loads too
                if (!ok) { s_fail = 1; break; }
            }
            if (push) {
                int pty, ptx;
                tile_of<PEER>(k, t_prev, pty, ptx);
                peer_push_tile<TC>(a, k, peer, push_src, pty, ptx, lane);
            }
            if (load) {
                int ty, txi;
                tile_of<PEER>(k, t, ty, txi);
                if (PEER && t < k.n_boundary) {
                    // peer-memory halo: the first / last tile row reads a row a neighbour
                    // stored and (through peer_push_tile) overwrites, in its memory, a row it
                    // may still be reading
                    int ok = 1;
                    if (lane == 0) {
                        if (peer.from_up && ty == 0 && !up_ok) {
                            s_pwait = 1;
                            ok = peer_flag_wait(peer.from_up, peer.need) ? 1 : 0;
                            s_pwait = 0;
                            up_ok = true;
                        }
                        if (ok && peer.from_down && ty == k.tiles_y - 1 && !down_ok) {
                            s_pwait = 1;
                            ok = peer_flag_wait(peer.from_down, peer.need) ? 1 : 0;
                            s_pwait = 0;
                            down_ok = true;
                        }
                    }
                    ok = __shfl_sync(kFull, ok, 0);
                    if (!ok) { s_fail = 1; break; }
                }
                if (lane == 0) {
                    const int c0 = txi * kTW, r0 = ty * kTR;
                    unsigned char *st = smem + (size_t)stage * SL::bytes;
                    mbar_arrive_expect_tx(&full_bar[stage], tx);
                    // tensor rows are counted from the top of the halo'd allocation
                    tma_load_2d(st + SL::oQ, &maps.Q, &full_bar[stage], c0 - 2, r0 - 1 + a.halo);
                    tma_load_2d(st + SL::oVol, &maps.vol, &full_bar[stage], c0, r0 + a.halo);
                    if (kKW) tma_load_2d(st + SL::oCoef, &maps.coef, &full_bar[stage], c0, r0);
                    if (use_i) tma_load_2d(st + SL::oI, &maps.I, &full_bar[stage], c0, r0 + a.halo);
                    if (use_h) tma_load_2d(st + SL::oH, &maps.H, &full_bar[stage], c0, r0 + a.halo);
                    if (use_t) tma_load_2d(st + SL::oT, &maps.T, &full_bar[stage], c0, r0 + a.halo);
                    tma_load_2d(st + SL::oW, &maps.w, &full_bar[stage], c0, r0);
                    tma_load_2d(st + SL::oGeo, &maps.geo, &full_bar[stage], c0, r0);
                    if (TFB_PK_RG_TMA) tma_load_2d(st + SL::oRG, &maps.rg, &full_bar[stage], 0, r0);
                }
            }
            if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
    } else {
        // ---------------- consumers: every lane advances NC = 2 cells TOGETHER -----------
        // (two adjacent columns where a warp takes a whole tile row, two tile rows where it
        // takes half a row): loads, then one straight-line block of arithmetic for both
        // cells, then the stores -- the dependent fp64 chains of the two cells interleave
        int stage = 0;
        unsigned phase = 0;
        constexpr int NC = (TC::cpl == 2) ? 2 : TC::rpw;
        static_assert(NC == 1 || NC == 2, "one or two cells in flight per lane");
        static_assert(TC::cpl == 1 || TC::rpw == 1, "two columns per lane: one row per warp and slot");
        // the tile coordinates advance without a division per tile (TileWalk, tfb_pk.cuh)
        TileWalk<PEER> walk;
        walk.init(k, (int)blockIdx.x);
        for (int t = blockIdx.x; t < k.n_tiles; t += gridDim.x) {
            if (!(PEER ? mbar_wait_peer(&full_bar[stage], phase, &s_pwait, &s_fail)
                       : mbar_wait(&full_bar[stage], phase))) { s_fail = 1; break; }
            const int ty = walk.ty, txi = walk.tx;
            const unsigned char *st = smem + (size_t)stage * SL::bytes;
            const int row0 = warp / TC::wpr;
            const int ct = (TC::cpl == 2) ? (((warp % TC::wpr) << 6) + lane * 2) : (((warp % TC::wpr) << 5) + lane);
            const int c0 = txi * kTW + ct, r0 = ty * kTR + row0;
            if (r0 < a.ny && c0 < a.nx) {
                PkIn x[NC];
                int rj[NC], cj[NC];
                bool live[NC];
                double coef[NC];
                if constexpr (TC::cpl == 2) {
                    const int e = row0 * kTW + ct;                          // element in the tile
                    const uint2 geo = *reinterpret_cast<const uint2 *>(st + SL::oGeo + e * 4);
                    const float2 w = *reinterpret_cast<const float2 *>(st + SL::oW + e * 4);
                    const double2 vol = *reinterpret_cast<const double2 *>(st + SL::oVol + e * 8);
                    double2 cf = make_double2(0.0, 0.0), Iv = cf, Hv = cf, Tv = cf;
                    if (kKW) cf = *reinterpret_cast<const double2 *>(st + SL::oCoef + e * 8);
                    if (use_i) Iv = *reinterpret_cast<const double2 *>(st + SL::oI + e * 8);
                    if (use_h) Hv = *reinterpret_cast<const double2 *>(st + SL::oH + e * 8);
                    if (use_t) Tv = *reinterpret_cast<const double2 *>(st + SL::oT + e * 8);
                    // Q window: halo rows row0 .. row0+2, halo cols ct .. ct+4
                    const double *q = reinterpret_cast<const double *>(st + SL::oQ) + row0 * kQW + ct;
                    const double2 a0 = *reinterpret_cast<const double2 *>(q);
                    const double2 a1 = *reinterpret_cast<const double2 *>(q + 2);
                    const double2 m0 = *reinterpret_cast<const double2 *>(q + kQW);
                    const double2 m1 = *reinterpret_cast<const double2 *>(q + kQW + 2);
                    const double2 b0 = *reinterpret_cast<const double2 *>(q + 2 * kQW);
                    const double2 b1 = *reinterpret_cast<const double2 *>(q + 2 * kQW + 2);
                    // columns: *0.y = c0-1, *1.x = c0, *1.y = c0+1, *E = c0+2
                    const double aE = q[4], mE = q[kQW + 4], bE = q[2 * kQW + 4];
                    // odd nx: the lane that holds the last column has no second cell.  Its inputs
                    // are the TMA unit's zero fill (code 0, vol 0, Q 0); with pixel area 0,
                    // width 1 and I = 1 it is a dry noflow cell that adds exactly 0 to every sum
                    // and counter, and its stores land in the row padding
                    live[0] = true; live[1] = (c0 + 1 < a.nx);
                    rj[0] = rj[1] = r0; cj[0] = c0; cj[1] = c0 + 1;
                    // children in the order SW, W, NW, N, NE, E, SE, S
                    x[0].g = geo.x; x[0].vol = vol.x; x[0].I = Iv.x; x[0].h = Hv.x; x[0].Ta = Tv.x;
                    x[0].w = (double)w.x; x[0].qself = m1.x; coef[0] = cf.x;
                    x[0].ch[0] = b0.y; x[0].ch[1] = m0.y; x[0].ch[2] = a0.y; x[0].ch[3] = a1.x;
                    x[0].ch[4] = a1.y; x[0].ch[5] = m1.y; x[0].ch[6] = b1.y; x[0].ch[7] = b1.x;
                    x[1].g = geo.y; x[1].vol = vol.y; x[1].I = live[1] ? Iv.y : 1.0; x[1].h = Hv.y;
                    x[1].Ta = Tv.y; x[1].w = live[1] ? (double)w.y : 1.0; x[1].qself = m1.y; coef[1] = cf.y;
                    x[1].ch[0] = b1.x; x[1].ch[1] = m1.x; x[1].ch[2] = a1.x; x[1].ch[3] = a1.y;
                    x[1].ch[4] = aE; x[1].ch[5] = mE; x[1].ch[6] = bE; x[1].ch[7] = b1.y;
                } else {
#pragma unroll
                    for (int j = 0; j < NC; ++j) {
                        const int row = row0 + j * TC::RP;
                        const int e = row * kTW + ct;
                        rj[j] = ty * kTR + row; cj[j] = c0;
                        live[j] = (rj[j] < a.ny);
                        x[j].g = *reinterpret_cast<const unsigned *>(st + SL::oGeo + e * 4);
                        x[j].w = live[j] ? (double)*reinterpret_cast<const float *>(st + SL::oW + e * 4) : 1.0;
                        x[j].vol = *reinterpret_cast<const double *>(st + SL::oVol + e * 8);
                        x[j].I = (use_i && live[j]) ? *reinterpret_cast<const double *>(st + SL::oI + e * 8) : 1.0;
                        x[j].h = use_h ? *reinterpret_cast<const double *>(st + SL::oH + e * 8) : 0.0;
                        x[j].Ta = use_t ? *reinterpret_cast<const double *>(st + SL::oT + e * 8) : 0.0;
                        coef[j] = kKW ? *reinterpret_cast<const double *>(st + SL::oCoef + e * 8) : 0.0;
                        // Q window: halo rows row .. row+2, halo cols ct+1 .. ct+3 (= c0-1 .. c0+1)
                        const double *q = reinterpret_cast<const double *>(st + SL::oQ) + row * kQW + ct + 1;
                        x[j].ch[0] = q[2 * kQW]; x[j].ch[1] = q[kQW]; x[j].ch[2] = q[0]; x[j].ch[3] = q[1];
                        x[j].ch[4] = q[2]; x[j].ch[5] = q[kQW + 2]; x[j].ch[6] = q[2 * kQW + 2];
                        x[j].ch[7] = q[2 * kQW + 1];
                        x[j].qself = q[kQW + 1];
                    }
                }
                // flow length of the cell's direction class, its reciprocal, pixel area of the row
                double ds[NC], ids[NC], da[NC];
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    const int rr = live[j] ? rj[j] : r0;
#if TFB_PK_RG_TMA
                    const double2 *rgp = reinterpret_cast<const double2 *>(st + SL::oRG) +
                                         (rr - ty * kTR) * (kRowGeom / 2);
                    const double2 g0 = rgp[0], g1 = rgp[1], g2 = rgp[2];
                    const double da_row = rgp[3].x;
#else
                    const double2 *rgp = reinterpret_cast<const double2 *>(p.row_geom + (int64_t)rr * kRowGeom);
                    const double2 g0 = __ldg(rgp), g1 = __ldg(rgp + 1), g2 = __ldg(rgp + 2);
                    const double da_row = __ldg(rgp + 3).x;
#endif
                    const unsigned code = x[j].g & 0xffu;
                    const bool dg = (code & 0x55u) != 0u, ns = (code & 0x88u) != 0u;
                    ds[j] = dg ? g0.y : (ns ? g1.x : g0.x);
                    ids[j] = dg ? g2.x : (ns ? g2.y : g1.y);
                    da[j] = live[j] ? da_row : 0.0;
                }
                // ---------------- the straight-line block --------------------------------
                PkVD o[NC];
                double u[NC], Qn[NC];
                bool bad[NC];
#pragma unroll
                for (int j = 0; j < NC; ++j) pk_vol_depth<SRC>(k, lut, x[j], ds[j], ids[j], da[j], acc, o[j]);
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    u[j] = 0.0; Qn[j] = 0.0; bad[j] = false;
                    if (kKW) bad[j] = pk_velocity(lut, x[j].g, o[j].d, x[j].w, coef[j], ids[j], rate_max, u[j], Qn[j]);
                }
                double d_pre[NC];
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    // update_edge_values :2960-2966 (pass A keeps the PRE-zero depth for the
                    // free-surface slope of pass B, which then zeroes it -- and for the
                    // neighbour slabs, whose halo rows the producer warp fills from it)
                    const bool nf = (x[j].g & 0xffu) == 0u;
                    d_pre[j] = o[j].d;
                    acc.s[S_VOL_EDGE] += nf ? o[j].v : 0.0;
                    o[j].v = nf ? 0.0 : o[j].v;
                    if (kKW) o[j].d = nf ? 0.0 : o[j].d;
                }
                // ---------------- rare tails ------------------------------------------------
                if (kKW) {
#pragma unroll
                    for (int j = 0; j < NC; ++j) {
                        if (bad[j]) pk_count_bad(acc, x[j].g, d_pre[j], u[j]);
                        if (live[j] && rj[j] + a.row0 == a.outlet_row && cj[j] == a.outlet_col)
                            pk_outlet(a, k.g, x[j].qself, u[j], d_pre[j], p.outlet_rough);
                    }
                }
                // ---------------- stores ----------------------------------------------------
                if constexpr (TC::cpl == 2) {
                    const int64_t i = (int64_t)r0 * a.pitch + c0;
                    if (kKW) {
                        *reinterpret_cast<double2 *>(a.u + i) = make_double2(u[0], u[1]);
                        *reinterpret_cast<double2 *>(a.Q_out + i) = make_double2(Qn[0], Qn[1]);
                    }
                    *reinterpret_cast<double2 *>(a.vol_out + i) = make_double2(o[0].v, o[1].v);
                    *reinterpret_cast<double2 *>(a.d + i) = make_double2(o[0].d, o[1].d);
                    if (use_i) *reinterpret_cast<double2 *>(a.I_out + i) = make_double2(o[0].I, o[1].I);
                    if (use_h) {
                        *reinterpret_cast<double2 *>(a.h_swe_out + i) = make_double2(o[0].h, o[1].h);
                        if (a.h_snow)
                            *reinterpret_cast<double2 *>(a.h_snow + i) =
                                make_double2(o[0].h * k.ratio, o[1].h * k.ratio);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < NC; ++j) {
                        if (!live[j]) continue;
                        const int64_t i = (int64_t)rj[j] * a.pitch + c0;
                        if (kKW) { a.u[i] = u[j]; a.Q_out[i] = Qn[j]; }
                        a.vol_out[i] = o[j].v;
                        a.d[i] = o[j].d;
                        if (use_i) a.I_out[i] = o[j].I;
                        if (use_h) {
                            a.h_swe_out[i] = o[j].h;
                            if (a.h_snow) a.h_snow[i] = o[j].h * k.ratio;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);            // slot may be refilled
            if (++stage == NS) { stage = 0; phase ^= 1u; }
            walk.advance(k, (int)gridDim.x);
        }
    }
    pk_finish<kConsumerWarps>(a, acc, rate_max, warp < kConsumerWarps, &s_fail, kKW ? 0 : 1, peer, sm,
                              tot, &is_last);
}

// ---------------------------------------------------------------------------
// pass B of the diffusive step: free-surface slope from the NEW depths of the
// cell and its D8 parent (channels_base.py:2583-2618), velocity, next discharge
// ---------------------------------------------------------------------------
template <class TC, bool PEER>
__global__ void __launch_bounds__(TC::threads, 1)
channels_packed_dwb_kernel(const tfb_channels_step_args a, const tfb_channels_packed p,
                           const PkConsts k, const __grid_constant__ PkMapsB maps, const PkPeer peer) {
    using SL = StageLayoutB<TC>;
    constexpr int kTR = TC::R, kTW = TC::W, kQW = TC::QW, kConsumerWarps = TC::warps, kThreads = TC::threads;
    constexpr int NS = SL::stages;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(16) double lut[256 * 4];
    __shared__ double sm[kConsumerWarps][kPartialStride];
    __shared__ double tot[kPartialStride];
    __shared__ __align__(8) uint64_t full_bar[NS], empty_bar[NS];
    __shared__ bool is_last;
    __shared__ volatile int s_fail;
    __shared__ volatile int s_pwait;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < p.n_angles * 4; i += kThreads) lut[i] = __ldg(p.angle_lut + i);
    if (tid < kConsumerWarps * kPartialStride) (&sm[0][0])[tid] = 0.0;
    if (tid == 0) {
        s_fail = 0;
        s_pwait = 0;
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
        // producer warp (see channels_packed_kernel): lane 0 issues the TMA loads, the whole
        // warp forwards finished boundary rows of the new discharge to the neighbours
        int stage = 0;
        unsigned phase = 0;
        bool up_ok = false, down_ok = false;
        const int t_end = k.n_tiles + (PEER ? NS * (int)gridDim.x : 0);
        for (int t = blockIdx.x; t < t_end; t += gridDim.x) {
            const int t_prev = t - NS * (int)gridDim.x;
            const bool push = PEER && t_prev >= 0 && t_prev < k.n_boundary;
            const bool load = t < k.n_tiles;
            if (load || push) {
                int ok = 1;
                if (lane == 0) ok = mbar_wait(&empty_bar[stage], phase ^ 1u) ? 1 : 0;
                ok = __shfl_sync(kFull, ok, 0);
                __syncwarp();                 // lane 0's acquire orders the other lanes' loads too
                if (!ok) { s_fail = 1; break; }
            }
            if (push) {
                int pty, ptx;
                tile_of<PEER>(k, t_prev, pty, ptx);
                peer_push_tile<TC>(a, k, peer, a.Q_out, pty, ptx, lane);
            }
            if (load) {
                int ty, txi;
                tile_of<PEER>(k, t, ty, txi);
                if (PEER && t < k.n_boundary) {
                    // the depth halo rows were stored by the neighbours' pass A of this step
                    int ok = 1;
                    if (lane == 0) {
                        if (peer.from_up && ty == 0 && !up_ok) {
                            s_pwait = 1;
                            ok = peer_flag_wait(peer.from_up, peer.need) ? 1 : 0;
                            s_pwait = 0;
                            up_ok = true;
                        }
                        if (ok && peer.from_down && ty == k.tiles_y - 1 && !down_ok) {
                            s_pwait = 1;
                            ok = peer_flag_wait(peer.from_down, peer.need) ? 1 : 0;
                            s_pwait = 0;
                            down_ok = true;
                        }
                    }
                    ok = __shfl_sync(kFull, ok, 0);
                    if (!ok) { s_fail = 1; break; }
                }
                if (lane == 0) {
                    const int c0 = txi * kTW, r0 = ty * kTR;
                    unsigned char *st = smem + (size_t)stage * SL::bytes;
                    mbar_arrive_expect_tx(&full_bar[stage], (unsigned)SL::tx_bytes);
                    tma_load_2d(st + SL::oD, &maps.D, &full_bar[stage], c0 - 2, r0 - 1 + a.halo);
                    tma_load_2d(st + SL::oN, &maps.N, &full_bar[stage], c0, r0);
                    tma_load_2d(st + SL::oS, &maps.S, &full_bar[stage], c0, r0);
                    tma_load_2d(st + SL::oW, &maps.w, &full_bar[stage], c0, r0);
                    tma_load_2d(st + SL::oGeo, &maps.geo, &full_bar[stage], c0, r0);
                    if (TFB_PK_RG_TMA) tma_load_2d(st + SL::oRG, &maps.rg, &full_bar[stage], 0, r0);
                }
            }
            if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
    } else {
        // consumers: one cell per lane, the NC = rpw tile rows a warp takes advance together
        int stage = 0;
        unsigned phase = 0;
        constexpr int NC = TC::rpw;
        static_assert(TC::cpl == 1 && (NC == 1 || NC == 2), "pass B: one cell per lane, <= 2 rows per warp");
        // the tile coordinates advance without a division per tile (TileWalk, tfb_pk.cuh)
        TileWalk<PEER> walk;
        walk.init(k, (int)blockIdx.x);
        for (int t = blockIdx.x; t < k.n_tiles; t += gridDim.x) {
            if (!(PEER ? mbar_wait_peer(&full_bar[stage], phase, &s_pwait, &s_fail)
                       : mbar_wait(&full_bar[stage], phase))) { s_fail = 1; break; }
            const int ty = walk.ty, txi = walk.tx;
            const unsigned char *st = smem + (size_t)stage * SL::bytes;
            const int row0 = warp / TC::wpr, ct = ((warp % TC::wpr) << 5) + lane;
            const int c0 = txi * kTW + ct, r0 = ty * kTR + row0;
            if (r0 < a.ny && c0 < a.nx) {
                unsigned g[NC];
                int rj[NC];
                bool live[NC], bad[NC];
                double d[NC], w[NC], umul[NC], ids[NC], u[NC], Qn[NC];
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    const int row = row0 + j * TC::RP;
                    const int e = row * kTW + ct;
                    rj[j] = ty * kTR + row;
                    live[j] = (rj[j] < a.ny);
                    g[j] = *reinterpret_cast<const unsigned *>(st + SL::oGeo + e * 4);
                    w[j] = live[j] ? (double)*reinterpret_cast<const float *>(st + SL::oW + e * 4) : 1.0;
                    const double Sb = (double)*reinterpret_cast<const float *>(st + SL::oS + e * 4);
                    const double invn = *reinterpret_cast<const double *>(st + SL::oN + e * 8);
                    const double *dc = reinterpret_cast<const double *>(st + SL::oD) + (row + 1) * kQW + ct + 2;
                    d[j] = dc[0];
                    const unsigned code = g[j] & 0xffu;
                    int dr, dcol;
                    code_offset(code, dr, dcol);
                    const double d_par = dc[dr * kQW + dcol];             // new depth of the D8 parent
#if TFB_PK_RG_TMA
                    const double2 *rgp = reinterpret_cast<const double2 *>(st + SL::oRG) +
                                         ((live[j] ? rj[j] : r0) - ty * kTR) * (kRowGeom / 2);
                    const double2 g1 = rgp[1], g2 = rgp[2];
#else
                    const double2 *rgp = reinterpret_cast<const double2 *>(
                        p.row_geom + (int64_t)(live[j] ? rj[j] : r0) * kRowGeom);
                    const double2 g1 = __ldg(rgp + 1), g2 = __ldg(rgp + 2);
#endif
                    const bool dg = (code & 0x55u) != 0u, ns = (code & 0x88u) != 0u;
                    ids[j] = dg ? g2.x : (ns ? g2.y : g1.y);
                    const double Sf = Sb + ((d[j] - d_par) * ids[j]);     // :2606-2607
                    umul[j] = sqrt_sel(fabs(Sf)) * invn;
                }
#pragma unroll
                for (int j = 0; j < NC; ++j)
                    bad[j] = pk_velocity(lut, g[j], d[j], w[j], umul[j], ids[j], rate_max, u[j], Qn[j]);
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    if (!live[j]) continue;
                    const int64_t i = (int64_t)rj[j] * a.pitch + c0;
                    if (bad[j]) pk_count_bad(acc, g[j], d[j], u[j]);
                    if (rj[j] + a.row0 == a.outlet_row && c0 == a.outlet_col)
                        pk_outlet(a, k.g, __ldg(a.Q_in + i), u[j], d[j], p.outlet_rough);
                    a.u[i] = u[j];
                    a.Q_out[i] = Qn[j];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == NS) { stage = 0; phase ^= 1u; }
            walk.advance(k, (int)gridDim.x);
        }
    }
    pk_finish<kConsumerWarps>(a, acc, rate_max, warp < kConsumerWarps, &s_fail, 2, peer, sm, tot,
                              &is_last);
}

}  // namespace tfb

using namespace tfb;

extern "C" int64_t tfb_sizeof_channels_packed(void) { return (int64_t)sizeof(tfb_channels_packed); }

extern "C" int tfb_channels_pack_geo32(const uint8_t *code, const uint16_t *angle_idx, int32_t ny,
                                       int32_t nx, int32_t halo, uint32_t *geo32, void *stream) {
    TFB_ARG_CHECK(code && geo32 && ny > 0 && nx > 0 && halo >= 0);
    dim3 b(32, 8), g((nx + 31) / 32, (ny + 7) / 8);
    pack_geo32_kernel<<<g, b, 0, (cudaStream_t)stream>>>(code, angle_idx, ny, nx, halo, geo32);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

extern "C" int tfb_channels_pack_rows(const float *dx, const float *dy, const float *dd,
                                      double sinu, const double *da_row, double da, int32_t ny,
                                      double *row_geom, void *stream) {
    TFB_ARG_CHECK(dx && dy && dd && row_geom && ny > 0);
    pack_rows_kernel<<<(ny + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dx, dy, dd, sinu, da_row,
                                                                       da, ny, row_geom);
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
    const void *ptr; int64_t rows, cols, pitch; int dtype; int box0, box1; int dev;
    bool operator==(const MapKey &o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && pitch == o.pitch && dtype == o.dtype &&
               box0 == o.box0 && box1 == o.box1 && dev == o.dev;
    }
};
struct MapEntry { MapKey key; CUtensorMap map; };
std::mutex g_map_mu;
std::vector<MapEntry> g_maps;
EncodeTiledFn g_encode = nullptr;

}  // namespace

int tfb::pk_get_map(const void *ptr, int64_t rows, int64_t cols, int64_t pitch, CUtensorMapDataType dtype,
            int esize, int box0, int box1, CUtensorMap *out) {
    int dev = 0;
    TFB_CUDA_CHECK(cudaGetDevice(&dev));
    const MapKey key{ptr, rows, cols, pitch, (int)dtype, box0, box1, dev};
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
    // the tensor is `cols` wide (cells beyond it read as zero fill), its rows `pitch` apart
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch * (cuuint64_t)esize};
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

namespace {

inline int get_map(const void *ptr, int64_t rows, int64_t cols, int64_t pitch, CUtensorMapDataType dtype,
                   int esize, int box0, int box1, CUtensorMap *out) {
    return tfb::pk_get_map(ptr, rows, cols, pitch, dtype, esize, box0, box1, out);
}

template <int SRC, int VAR, class TC, bool PEER>
int launch_main_(const tfb_channels_step_args &a, const tfb_channels_packed &p, PkConsts k,
                 cudaStream_t st, const PkPeer &peer) {
    using SL = StageLayout<SRC, VAR, TC>;
    constexpr int kTR = TC::R, kQR = TC::QR, kTW = TC::W, kQW = TC::QW, kThreads = TC::threads;
    k.tiles_x = (a.nx + kTW - 1) / kTW;
    k.tiles_y = (a.ny + kTR - 1) / kTR;
    k.n_tiles = k.tiles_x * k.tiles_y;
    k.n_boundary = k.tiles_x * (k.tiles_y > 1 ? 2 : 1);
    // serpentine: an odd kinematic step starts on the rows the even one wrote last (still in L2)
    k.reverse = (!PEER && VAR == PK_KW && (a.flags & TFB_CH_REVERSE)) ? 1 : 0;
    static_assert(SL::stages >= 2, "ring needs at least two slots");
    PkMaps m;
    const int64_t rows_h = (int64_t)a.ny + 2 * a.halo;
    const size_t hb = (size_t)a.halo * a.pitch;            // elements between allocation and row 0
    const CUtensorMapDataType F64 = CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    int rc;
    if ((rc = get_map(a.Q_in - hb, rows_h, a.nx, a.pitch, F64, 8, kQW, kQR, &m.Q))) return rc;
    if ((rc = get_map(a.vol - hb, rows_h, a.nx, a.pitch, F64, 8, kTW, kTR, &m.vol))) return rc;
    m.coef = m.I = m.H = m.T = m.vol;
    if (VAR == PK_KW)
        if ((rc = get_map(p.coef, a.ny, a.nx, a.pitch, F64, 8, kTW, kTR, &m.coef))) return rc;
    const bool use_h = (SRC == PK_SRC_FULL) && (k.flags & TFB_CH_SRC_SNOW);
    const bool use_i = (SRC == PK_SRC_GA) || ((SRC == PK_SRC_FULL) && (k.flags & TFB_CH_SRC_INFIL));
    const bool use_t = (SRC == PK_SRC_FULL) && k.tair_grid;
    if (use_i) if ((rc = get_map(a.I - hb, rows_h, a.nx, a.pitch, F64, 8, kTW, kTR, &m.I))) return rc;
    if (use_h) if ((rc = get_map(a.h_swe - hb, rows_h, a.nx, a.pitch, F64, 8, kTW, kTR, &m.H))) return rc;
    if (use_t) if ((rc = get_map(a.T_air.ptr - hb, rows_h, a.nx, a.pitch, F64, 8, kTW, kTR, &m.T))) return rc;
    if ((rc = get_map(p.width32, a.ny, a.nx, a.pitch, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, kTW, kTR, &m.w))) return rc;
    if ((rc = get_map(p.geo32, a.ny, a.nx, a.pitch, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, kTW, kTR, &m.geo))) return rc;
    m.rg = m.geo;
    if (TFB_PK_RG_TMA)
        if ((rc = get_map(p.row_geom, a.ny, kRowGeom, kRowGeom, F64, 8, kRowGeom, kTR, &m.rg))) return rc;
    const int smem = SL::stages * SL::bytes;
    static bool attr_set[64] = {false};
    int dev = 0;
    TFB_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        TFB_CUDA_CHECK(cudaFuncSetAttribute(channels_packed_kernel<SRC, VAR, TC, PEER>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (dev < 64) attr_set[dev] = true;
    }
    int64_t grid = sm_count();
    if (grid > k.n_tiles) grid = k.n_tiles;
    channels_packed_kernel<SRC, VAR, TC, PEER><<<(int)grid, kThreads, smem, st>>>(a, p, k, m, peer);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

template <int SRC, int VAR, class TC>
int launch_main(const tfb_channels_step_args &a, const tfb_channels_packed &p, const PkConsts &k,
                cudaStream_t st) {
    return launch_main_<SRC, VAR, TC, false>(a, p, k, st, PkPeer{});
}
template <int SRC, class TC, int VAR = PK_KW>
int launch_main_peer(const tfb_channels_step_args &a, const tfb_channels_packed &p, const PkConsts &k,
                     cudaStream_t st, const PkPeer &peer) {
    return launch_main_<SRC, VAR, TC, true>(a, p, k, st, peer);
}

template <class TC, bool PEER>
int launch_dwb(const tfb_channels_step_args &a, const tfb_channels_packed &p, PkConsts k,
               cudaStream_t st, const PkPeer &peer) {
    using SL = StageLayoutB<TC>;
    constexpr int kTR = TC::R, kQR = TC::QR, kTW = TC::W, kQW = TC::QW, kThreads = TC::threads;
    k.tiles_x = (a.nx + kTW - 1) / kTW;
    k.tiles_y = (a.ny + kTR - 1) / kTR;
    k.n_tiles = k.tiles_x * k.tiles_y;
    k.n_boundary = k.tiles_x * (k.tiles_y > 1 ? 2 : 1);
    // pass A walks top-down, pass B bottom-up: B starts on the depths A wrote last, the next
    // step's A on the discharges B wrote last -- both while those rows are still in L2
    k.reverse = (!PEER && (a.flags & TFB_CH_REVERSE)) ? 1 : 0;      // (the caller sets it for pass B)
    PkMapsB m;
    const int64_t rows_h = (int64_t)a.ny + 2 * a.halo;
    const size_t hb = (size_t)a.halo * a.pitch;
    int rc;
    if ((rc = get_map(a.d - hb, rows_h, a.nx, a.pitch, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, kQW, kQR, &m.D))) return rc;
    if ((rc = get_map(p.inv_rough, a.ny, a.nx, a.pitch, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, kTW, kTR, &m.N))) return rc;
    if ((rc = get_map(p.S32, a.ny, a.nx, a.pitch, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, kTW, kTR, &m.S))) return rc;
    if ((rc = get_map(p.width32, a.ny, a.nx, a.pitch, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, kTW, kTR, &m.w))) return rc;
    if ((rc = get_map(p.geo32, a.ny, a.nx, a.pitch, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, kTW, kTR, &m.geo))) return rc;
    m.rg = m.geo;
    if (TFB_PK_RG_TMA)
        if ((rc = get_map(p.row_geom, a.ny, kRowGeom, kRowGeom, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, kRowGeom, kTR, &m.rg))) return rc;
    const int smem = SL::stages * SL::bytes;
    static bool attr_set[64] = {false};
    int dev = 0;
    TFB_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        TFB_CUDA_CHECK(cudaFuncSetAttribute(channels_packed_dwb_kernel<TC, PEER>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (dev < 64) attr_set[dev] = true;
    }
    int64_t grid = sm_count();
    if (grid > k.n_tiles) grid = k.n_tiles;
    channels_packed_dwb_kernel<TC, PEER><<<(int)grid, kThreads, smem, st>>>(a, p, k, m, peer);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

// argument checks and constants shared by the three entry points
int prepare(const tfb_channels_step_args *ap, tfb_channels_step_args &a,
            const tfb_channels_packed *pp, bool diffusive, PkConsts &k, int &src_mode) {
    TFB_ARG_CHECK(ap != nullptr && pp != nullptr);
    a = *ap;
    if (a.pitch == 0) a.pitch = a.nx;                // dense rows
    const tfb_channels_packed &p = *pp;
    // TMA: 16-byte row strides for the 8-byte and the 4-byte streams alike
    TFB_ARG_CHECK(a.ny > 0 && a.nx > 0 && a.halo >= 0 && a.pitch >= a.nx && (a.pitch % 4) == 0);
    TFB_ARG_CHECK(p.geo32 && p.width32 && p.angle_lut && p.row_geom && p.n_angles >= 1 &&
                  p.n_angles <= 256);
    TFB_ARG_CHECK(a.vol && a.vol_out && a.Q_in && a.Q_out && a.d && a.u);
    TFB_ARG_CHECK(a.Q_in != a.Q_out && a.scalars && a.partials && a.ticket && a.dt > 0.0);
    // eligibility: Manning, lean outputs, uniform forcing
    const uint32_t src = a.flags & (TFB_CH_SRC_MET | TFB_CH_SRC_SNOW | TFB_CH_SRC_EVAP |
                                    TFB_CH_SRC_INFIL);
    TFB_ARG_CHECK(!(a.flags & (TFB_CH_LAW_OF_WALL | TFB_CH_DIAG | TFB_CH_WRITE_R | TFB_CH_FLOOD |
                               TFB_CH_ATTENUATE)));
    TFB_ARG_CHECK(((a.flags & TFB_CH_DIFFUSIVE) != 0) == diffusive);
    if (diffusive) TFB_ARG_CHECK(p.inv_rough != nullptr && p.S32 != nullptr);
    else TFB_ARG_CHECK(p.coef != nullptr);
    TFB_ARG_CHECK(!a.sinu.ptr);      // the pixel area comes from row_geom (scalar or per row)
    TFB_ARG_CHECK(!a.P_rain.ptr && !a.SM.ptr && !a.GW.ptr && !a.MR.ptr && !a.ET.ptr && !a.IN.ptr);
    auto al16 = [](const void *q) { return (((uintptr_t)q) & 15u) == 0; };
    TFB_ARG_CHECK(al16(a.vol) && al16(a.vol_out) && al16(a.Q_in) && al16(a.Q_out) && al16(a.d) &&
                  al16(a.u) && al16(p.coef) && al16(p.inv_rough) && al16(p.row_geom) &&
                  ((((uintptr_t)p.geo32) & 7u) == 0) && ((((uintptr_t)p.width32) & 7u) == 0) &&
                  ((((uintptr_t)p.S32) & 7u) == 0));
    memset(&k, 0, sizeof(k));
    k.g = 9.81;
    k.dt = a.dt;
    k.da = a.da.val;
    k.Psum = ((a.P_rain.val + a.SM.val) + a.GW.val) + a.MR.val;
    k.P_total = a.P_rain.val + a.SM.val;
    k.tiles_x = 0;
    k.n_tiles = 0;                                   // set per tile configuration by the launcher
    k.flags = src;
    if (src == 0) {
        TFB_ARG_CHECK((a.is_R_grid != 0) == (a.da.ptr != nullptr));
        k.Runi = k.Psum - (0.0 + a.IN.val);
        k.r_grid = a.is_R_grid ? 1 : 0;
        src_mode = PK_SRC_NONE;
        return TFB_OK;
    }
    TFB_ARG_CHECK(a.is_R_grid == 1);
    if (src & TFB_CH_SRC_INFIL) {
        TFB_ARG_CHECK(a.I && a.I_out && al16(a.I) && al16(a.I_out));
        TFB_ARG_CHECK(!a.Ks.ptr && !a.Ki.ptr && !a.qs.ptr && !a.qi.ptr && !a.G.ptr &&
                      !a.h_table.ptr && !a.elev.ptr);
        TFB_ARG_CHECK(!(a.h_table.val >= a.elev.val));          // infil_base.py:697-701
        k.Ks = a.Ks.val;
        k.gkq = (a.G.val * (a.Ks.val - a.Ki.val)) * (a.qs.val - a.qi.val);
        k.low_rain = (k.P_total < k.Ks) ? 1 : 0;
    }
    if (src == TFB_CH_SRC_INFIL) {
        src_mode = PK_SRC_GA;
        return TFB_OK;
    }
    // general fused sources: every parameter a scalar, T_air optionally a grid
    src_mode = PK_SRC_FULL;
    TFB_ARG_CHECK(!a.c0.ptr && !a.T0.ptr && !a.T_surf.ptr && !a.Qn_SW.ptr && !a.Qn_LW.ptr);
    TFB_ARG_CHECK(al16(a.T_air.ptr));
    if (src & TFB_CH_SRC_SNOW)
        TFB_ARG_CHECK(a.h_swe && a.h_swe_out && al16(a.h_swe) && al16(a.h_swe_out) && al16(a.h_snow));
    k.tair_grid = a.T_air.ptr ? 1 : 0;
    k.P = a.P_rain.val; k.SMv = a.SM.val; k.GW = a.GW.val; k.MR = a.MR.val;
    k.ETv = a.ET.val; k.INv = a.IN.val; k.T_air = a.T_air.val; k.T_rs = a.T_rain_snow;
    k.c0_864 = a.c0.val / 8.64E7;
    k.T0 = a.T0.val;
    k.ratio = a.rho_H2O / a.rho_snow;
    k.Qc = a.K_soil * (a.T_surf.val - a.T_soil_x) / a.soil_x;     // evap_base.py:326-353
    k.Qnet = a.Qn_SW.val + a.Qn_LW.val;
    k.alpha = a.alpha;
    return TFB_OK;
}

}  // namespace

// tile configurations (rows, cells per lane), chosen from measurements on B200: variants
// that fit 64 registers run two warps per tile row (31 warps per SM), the register-
// hungry fused-source variants one warp per row with 128 registers
#ifndef TFB_PK_GA_ROWS
#define TFB_PK_GA_ROWS 15
#endif
#ifndef TFB_PK_GA_CPL
#define TFB_PK_GA_CPL 2
#endif
#ifndef TFB_PK_GA_W
#define TFB_PK_GA_W 64
#endif
#ifndef TFB_PK_NONE_ROWS
#define TFB_PK_NONE_ROWS 24       // KW without sources: 2 cells per lane, 24-row tiles (25 warps, 72 registers,
#endif                            // 3 ring slots of 51 KB).  Same-box A/B, round 2 (profiles/r02/ab_tiles*.log): 1 cell per
                                  // lane x 2 rows per warp 0.2172 ms; 2 cells per lane: 15 rows 0.2094, 17 rows
                                  // 0.2066, 20 rows 0.1956, 22 rows 0.1915, 24 rows 0.1869, 26 rows (3 slots) 0.2126
#ifndef TFB_PK_NONEA_ROWS
#define TFB_PK_NONEA_ROWS 20      // DW pass A without sources: 15 rows 0.2903 ms (step), 20 rows 0.2824, 24 rows 0.2824
#endif
#ifndef TFB_PK_NONE_W
#define TFB_PK_NONE_W 64
#endif
#ifndef TFB_PK_NONE_RPW
#define TFB_PK_NONE_RPW 1
#endif
#ifndef TFB_PK_GA_RPW
#define TFB_PK_GA_RPW 1
#endif
#ifndef TFB_PK_FULL_RPW
#define TFB_PK_FULL_RPW 1
#endif
#ifndef TFB_PK_DWB_RPW
#define TFB_PK_DWB_RPW 2
#endif
#ifndef TFB_PK_NONE_CPL
#define TFB_PK_NONE_CPL 2
#endif
typedef TileCfg<TFB_PK_NONE_ROWS, TFB_PK_NONE_CPL, TFB_PK_NONE_W, TFB_PK_NONE_RPW> CfgNone;
typedef TileCfg<TFB_PK_NONEA_ROWS, TFB_PK_NONE_CPL, TFB_PK_NONE_W, TFB_PK_NONE_RPW> CfgNoneA;
typedef TileCfg<TFB_PK_GA_ROWS, TFB_PK_GA_CPL, TFB_PK_GA_W, TFB_PK_GA_RPW> CfgGA;
#ifndef TFB_PK_FULL_ROWS
#define TFB_PK_FULL_ROWS 15
#endif
#ifndef TFB_PK_FULL_CPL
#define TFB_PK_FULL_CPL 2
#endif
typedef TileCfg<TFB_PK_FULL_ROWS, TFB_PK_FULL_CPL, 64, TFB_PK_FULL_RPW> CfgFull;
typedef TileCfg<15, 1, 64, TFB_PK_DWB_RPW> CfgDwB;

extern "C" int tfb_channels_step_packed(const tfb_channels_step_args *ap,
                                        const tfb_channels_packed *pp, void *stream) {
    PkConsts k;
    int mode = 0;
    tfb_channels_step_args an;
    const int rc = prepare(ap, an, pp, false, k, mode);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == PK_SRC_NONE) return launch_main<PK_SRC_NONE, PK_KW, CfgNone>(an, *pp, k, st);
    if (mode == PK_SRC_GA) return launch_main<PK_SRC_GA, PK_KW, CfgGA>(an, *pp, k, st);
    return launch_main<PK_SRC_FULL, PK_KW, CfgFull>(an, *pp, k, st);
}

extern "C" int tfb_channels_step_packed_dw_c(const tfb_channels_step_args *ap,
                                             const tfb_channels_packed *pp, void *stream) {
    TFB_ARG_CHECK(ap && pp && ap->d && pp->n_noflow >= 0 && (pp->n_noflow == 0 || pp->noflow_idx));
    if (pp->n_noflow == 0) return TFB_OK;
    zero_at_kernel<<<(unsigned)((pp->n_noflow + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        ap->d, pp->noflow_idx, pp->n_noflow);
    if ((ap->flags & TFB_CH_FLOOD) && ap->d_flood)       // update_edge_values :2968-2970
        zero_at_kernel<<<(unsigned)((pp->n_noflow + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            ap->d_flood, pp->noflow_idx, pp->n_noflow);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

// ---- peer-memory halo -----------------------------------------------------------------
namespace {
typedef CUresult (*StreamWaitFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamWaitFn g_wait32 = nullptr;

int stream_wait_geq(cudaStream_t st, const uint32_t *addr, uint32_t value) {
    if (!g_wait32) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        TFB_CUDA_CHECK(cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) {
            tfb::set_error("cuStreamWaitValue32 not available from the driver");
            return TFB_ERR_CUDA;
        }
        g_wait32 = (StreamWaitFn)fn;
    }
    const CUresult rc = g_wait32((CUstream)st, (CUdeviceptr)(uintptr_t)addr, value,
                                 CU_STREAM_WAIT_VALUE_GEQ);
    if (rc != CUDA_SUCCESS) {
        tfb::set_error("cuStreamWaitValue32 failed (%d)", (int)rc);
        return TFB_ERR_CUDA;
    }
    return TFB_OK;
}
}  // namespace

extern "C" int64_t tfb_sizeof_peer_halo(void) { return (int64_t)sizeof(tfb_peer_halo); }

extern "C" int tfb_peer_alloc(int64_t bytes, void **ptr, uint8_t *handle64) {
    TFB_ARG_CHECK(bytes > 0 && ptr && handle64);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    TFB_CUDA_CHECK(cudaMalloc(ptr, (size_t)bytes));
    TFB_CUDA_CHECK(cudaMemset(*ptr, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    TFB_CUDA_CHECK(cudaIpcGetMemHandle(&h, *ptr));
    memcpy(handle64, &h, 64);
    return TFB_OK;
}

extern "C" int tfb_peer_open(const uint8_t *handle64, void **ptr) {
    TFB_ARG_CHECK(handle64 && ptr);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    TFB_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TFB_OK;
}

extern "C" int tfb_peer_close(void *ptr) {
    TFB_ARG_CHECK(ptr != nullptr);
    TFB_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
    return TFB_OK;
}

extern "C" int tfb_peer_free(void *ptr) {
    TFB_ARG_CHECK(ptr != nullptr);
    TFB_CUDA_CHECK(cudaFree(ptr));
    return TFB_OK;
}

namespace {
// kernel-side view of a tfb_peer_halo: publish step + 1 when my boundary tiles are done,
// wait (in the boundary tiles' producers, or on the stream) for the neighbours' `step`
int make_peer(const tfb_peer_halo *ph, cudaStream_t st, PkPeer &peer) {
    peer.up_Q = ph->up_Q; peer.down_Q = ph->down_Q;
    peer.up_flag = ph->up_flag; peer.down_flag = ph->down_flag;
    peer.value = ph->step + 1u;
    if (!getenv("TFB_P2P_STREAM_WAIT")) {
        // the producer warps of the boundary tiles poll the flags themselves: no stream
        // memory operation (several microseconds each) between consecutive launches
        peer.from_up = ph->from_up; peer.from_down = ph->from_down;
        peer.need = ph->step;
        return TFB_OK;
    }
    int rc;
    peer.from_up = peer.from_down = nullptr;
    peer.need = 0u;
    if (ph->from_up && (rc = stream_wait_geq(st, ph->from_up, ph->step))) return rc;
    if (ph->from_down && (rc = stream_wait_geq(st, ph->from_down, ph->step))) return rc;
    return TFB_OK;
}
}  // namespace

extern "C" int tfb_channels_step_packed_p2p(const tfb_channels_step_args *ap,
                                            const tfb_channels_packed *pp,
                                            const tfb_peer_halo *ph, void *stream) {
    TFB_ARG_CHECK(ph != nullptr);
    PkConsts k;
    int mode = 0;
    tfb_channels_step_args an;
    int rc = prepare(ap, an, pp, false, k, mode);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // both neighbours must have finished the previous step: their halo rows are in
    // my memory, and they no longer read the halo rows this step overwrites
    PkPeer peer;
    if ((rc = make_peer(ph, st, peer))) return rc;
    if (mode == PK_SRC_NONE) return launch_main_peer<PK_SRC_NONE, CfgNone>(an, *pp, k, st, peer);
    if (mode == PK_SRC_GA) return launch_main_peer<PK_SRC_GA, CfgGA>(an, *pp, k, st, peer);
    return launch_main_peer<PK_SRC_FULL, CfgFull>(an, *pp, k, st, peer);
}

extern "C" int tfb_channels_step_packed_dw_a_p2p(const tfb_channels_step_args *ap,
                                                 const tfb_channels_packed *pp,
                                                 const tfb_peer_halo *ph, void *stream) {
    TFB_ARG_CHECK(ph != nullptr);
    PkConsts k;
    int mode = 0;
    tfb_channels_step_args an;
    int rc = prepare(ap, an, pp, true, k, mode);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    PkPeer peer;
    if ((rc = make_peer(ph, st, peer))) return rc;
    if (mode == PK_SRC_NONE) return launch_main_peer<PK_SRC_NONE, CfgNoneA, PK_DWA>(an, *pp, k, st, peer);
    if (mode == PK_SRC_GA) return launch_main_peer<PK_SRC_GA, CfgGA, PK_DWA>(an, *pp, k, st, peer);
    return launch_main_peer<PK_SRC_FULL, CfgFull, PK_DWA>(an, *pp, k, st, peer);
}

extern "C" int tfb_channels_step_packed_dw_a(const tfb_channels_step_args *ap,
                                             const tfb_channels_packed *pp, void *stream) {
    PkConsts k;
    int mode = 0;
    tfb_channels_step_args an;
    const int rc = prepare(ap, an, pp, true, k, mode);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == PK_SRC_NONE) return launch_main<PK_SRC_NONE, PK_DWA, CfgNoneA>(an, *pp, k, st);
    if (mode == PK_SRC_GA) return launch_main<PK_SRC_GA, PK_DWA, CfgGA>(an, *pp, k, st);
    return launch_main<PK_SRC_FULL, PK_DWA, CfgFull>(an, *pp, k, st);
}

extern "C" int tfb_channels_step_packed_dw_b(const tfb_channels_step_args *ap,
                                             const tfb_channels_packed *pp, void *stream) {
    PkConsts k;
    int mode = 0;
    tfb_channels_step_args an;
    const int rc = prepare(ap, an, pp, true, k, mode);
    if (rc) return rc;
    return launch_dwb<CfgDwB, false>(an, *pp, k, (cudaStream_t)stream, PkPeer{});
}

extern "C" int tfb_channels_step_packed_dw_b_p2p(const tfb_channels_step_args *ap,
                                                 const tfb_channels_packed *pp,
                                                 const tfb_peer_halo *ph, void *stream) {
    TFB_ARG_CHECK(ph != nullptr);
    PkConsts k;
    int mode = 0;
    tfb_channels_step_args an;
    int rc = prepare(ap, an, pp, true, k, mode);
    if (rc) return rc;
    PkPeer peer;
    if ((rc = make_peer(ph, (cudaStream_t)stream, peer))) return rc;
    return launch_dwb<CfgDwB, true>(an, *pp, k, (cudaStream_t)stream, peer);
}
