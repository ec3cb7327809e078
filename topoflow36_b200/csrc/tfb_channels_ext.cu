// Extended variants of the packed (TMA ring) routing step: everything a COUPLED run hands
// the channels component that the lean production kernels of tfb_channels_packed.cu leave
// out --
//   * any of the six forcing inputs P_rain, SM, GW, MR, ET, IN as a fresh (ny, nx) grid
//     (update_R, channels_base.py:1548-1655, incl. the in-place masking of the provider's
//     ET grid :1625-1627 and the R grid itself),
//   * the law of the wall (law_of_the_wall :4005-4063) next to Manning's formula,
//   * the seven diagnostic grids materialised every step (update_trapezoid_Rh :2643,
//     update_shear_stress / _speed :2620-2641, update_friction_factor :2714,
//     update_froude_number :2867, + S_free for the diffusive wave :2583),
//   * more than 256 distinct bank angles (16-bit index into a table in L2), any nx,
//   * ATTENUATE: the per-cell linear reservoir in front of the channel (:2077-2084), one
//     more fp64 state stream in and out,
//   * FLOOD_OPTION: rectangle-over-trapezoid overbank flow (update_flood_discharge :1701,
//     update_discharge :1791, update_channel_volume :2149, update_flood_volume :2158,
//     update_flood_depth :2393): a bankfull-volume stream, flood depth / flood volume grids
//     out, and -- diffusive wave -- the flood depths of a cell and its D8 parent as a second
//     halo tile in pass B (the free surface is the TOTAL depth :2593-2596)
// -- for the kinematic wave (one launch) and the diffusive wave (pass A: volumes + new
// depths, pass B: free-surface slope, velocity, discharge, diagnostics; the noflow depths are
// zeroed by tfb_channels_step_packed_dw_c as in the lean path).
//
// Same machinery as the lean kernels (one persistent CTA per SM, a producer lane streaming
// 64 x 15-cell tiles of every input stream into a shared-memory ring with TMA, consumer
// warps reading the D8 children from the shared Q (or depth) halo tile, deterministic
// reductions), but the ring slot is laid out at RUN time -- an optional forcing grid is one
// more fp64 stream (+8 B per cell of HBM traffic, counted in the roofline) -- and the
// arithmetic keeps the reference's operation order with IEEE divisions (this TU is compiled
// with -fmad=false like tfb_channels.cu), so results equal the generic kernel's except
// for x^(2/3) (pow23 instead of libdevice cbrt, both within 1-2 ulp of numpy).
#include "tfb_pk.cuh"

#include <string.h>

namespace tfb {

enum { EXT_KW = 0, EXT_DWA = 1, EXT_DWB = 2 };
enum { F_P = 0, F_SM, F_GW, F_MR, F_ET, F_IN, F_N };

struct ExtConsts {
    double g, dt, aval, kappa, law_const, rho;
    double fval[F_N];          // scalar value of each forcing input (no stream)
    int foff[F_N];             // ring-slot offset of its stream, -1 = scalar
    int et_grid, r_grid, write_R;
    int oQ, oVol, oGeo, oW, oS, oRough, stage_bytes, n_stages;
    unsigned tx_bytes;
    int n_tiles, tiles_x, tiles_y;
    int d00_visible;           // pass B: global cell (0, 0) lies within this slab's rows + halo
    long long d00_index;       //         its offset from the first owned row
    int atten, flood;          // ATTENUATE, FLOOD_OPTION
    int oVs, oVb, oDf;         // ring-slot offsets: vol_stored, vol_bankfull (-1 = scalar), d_flood halo tile
    int w64, s64, oSinu;       // width / bed slope streams are fp64 grids; sinuosity grid stream (-1 = none)
    int oRG;                   // the tile rows' geometry (8 doubles per row), one more TMA box
    int reverse;               // walk the tiles bottom-up (TFB_CH_REVERSE; pass B always): see tfb.h
    double t_drain, vb_val, width_ratio, flood_n;
};
struct ExtMaps { CUtensorMap Q, vol, geo, w, S, rough, F[F_N], vs, vb, df, sinu, rg; };   // Q: discharge, or depth (pass B)

template <int VAR, bool LOW, bool DIAG>
__device__ __forceinline__ void ext_cell(const tfb_channels_step_args &a, const tfb_channels_packed &p,
                                         const ExtConsts &k, const unsigned char *st, int QW,
                                         int TW, int row, int ct, int r, int c, Acc &acc,
                                         float &rate_max) {
    const int e = row * TW + ct;
    const unsigned geo = *reinterpret_cast<const unsigned *>(st + k.oGeo + e * 4);
    const unsigned code = geo & 0xffu, cmask = (geo >> 8) & 0xffu, ai = geo >> 16;
    const bool noflow = (code == 0u);
    const double w = k.w64 ? *reinterpret_cast<const double *>(st + k.oW + e * 8)
                           : (double)*reinterpret_cast<const float *>(st + k.oW + e * 4);
    // {tan, cos} of the cell's bank angle from the table in L2 (a copy in shared memory for
    // <= 256 angles measured 0.5-2 % SLOWER: the kernel is bound by issue slots and registers)
    const double2 tc = __ldg(reinterpret_cast<const double2 *>(p.lut_tc) + ai);
    // the row's geometry from the ring slot: ds_ew, ds_diag, ds_ns, ., ., ., da
    const double *rg = reinterpret_cast<const double *>(st + k.oRG) + row * kRowGeom;
    const bool dg = (code & 0x55u) != 0u, ns = (code & 0x88u) != 0u;
    double ds = rg[dg ? 1 : (ns ? 2 : 0)];
    if (k.oSinu >= 0) ds = *reinterpret_cast<const double *>(st + k.oSinu + e * 8) * ds;   // :1242
    const int64_t i = (int64_t)r * a.pitch + c;
    const double dt = k.dt;
    double v = 0.0, d = 0.0, qself = 0.0;
    double df = 0.0, vf = 0.0;                   // FLOOD_OPTION: flood depth, volume above bankfull
    if (VAR != EXT_DWB) {
        // ---- update_R :1548-1655, update_R_integral :1657, update_flow_volume :1951 --------
        const double da = rg[6];
        const double vol = *reinterpret_cast<const double *>(st + k.oVol + e * 8);
        double f[F_N];
#pragma unroll
        for (int j = 0; j < F_N; ++j)
            f[j] = (k.foff[j] >= 0) ? *reinterpret_cast<const double *>(st + k.foff[j] + e * 8)
                                    : k.fval[j];
        double ET = f[F_ET];
        if (k.et_grid) {
            if (vol < ((ET * da) * dt)) {                            // :1625-1627, in place
                ET = 0.0;
                if (a.ET_rw) a.ET_rw[i] = 0.0;
            }
        } else {
            ET = 0.0;                                                // :1629 a scalar ET is dropped
        }
        const double R = (((f[F_P] + f[F_SM]) + f[F_GW]) + f[F_MR]) - (ET + f[F_IN]);   // :1649
        double Rv = (R * da) * dt;
        if (k.r_grid) acc.s[S_VOL_R] += Rv;                          // :1666-1670
        if (k.write_R) a.R[i] = R;
        if (k.atten) {
            // ATTENUATE :2077-2084: the runoff fills a linear reservoir that drains into the channel
            double vs = *reinterpret_cast<const double *>(st + k.oVs + e * 8) + Rv;
            const double Q_sides = div_sel(vs, k.t_drain);
            Rv = np_min(Q_sides * dt, vs);
            vs = vs - Rv;
            a.vol_stored_out[i] = np_max0(vs);
        }
        v = vol + Rv;                                                // :2086
        // children in the reference's scatter order SW, W, NW, N, NE, E, SE, S :2116-2131
        const double *q = reinterpret_cast<const double *>(st + k.oQ) + row * QW + ct + 1;
        add_if(v, cmask, 1u, dt * q[2 * QW]);
        add_if(v, cmask, 2u, dt * q[QW]);
        add_if(v, cmask, 4u, dt * q[0]);
        add_if(v, cmask, 8u, dt * q[1]);
        add_if(v, cmask, 16u, dt * q[2]);
        add_if(v, cmask, 32u, dt * q[QW + 2]);
        add_if(v, cmask, 64u, dt * q[2 * QW + 2]);
        add_if(v, cmask, 128u, dt * q[2 * QW + 1]);
        qself = q[QW + 1];
        v -= (qself * dt);                                           // :2138
        v = np_max0(v);                                              // :2145
        double vch = v;
        if (k.flood) {
            // update_channel_volume :2149, update_flood_volume :2158-2185, update_flood_depth
            // :2393-2448: the channel depth below comes from min(vol, bankfull volume)
            const double vb = (k.oVb >= 0) ? *reinterpret_cast<const double *>(st + k.oVb + e * 8)
                                           : k.vb_val;
            vf = np_max0(v - vb);
            vch = np_min(v, vb);
            df = np_max0(div_sel(vf, (k.width_ratio * w) * ds));
        }
        // ---- update_channel_depth :2299-2330, :2379 ----------------------------------------
        // both forms, straight-line (exactly rounded Newton division / sqrt of tfb_common.cuh,
        // IEEE special cases by selects); the bank angle picks one
        const double denom = 2.0 * tc.x;
        double arg = div_sel((2.0 * denom) * vch, ds);
        arg += w * w;
        const double d_trap = div_sel(sqrt_sel(arg) - w, denom);
        const double d_rect = div_sel(vch, w * ds);
        d = np_max0((tc.x == 0.0) ? d_rect : d_trap);
    }
    if (VAR == EXT_DWA) {
        // pass A ends here: the PRE-zero depth feeds the free-surface slope of pass B
        if (noflow) { acc.s[S_VOL_EDGE] += v; v = 0.0; }             // :2960-2964
        a.vol_out[i] = v;
        a.d[i] = d;
        if (k.flood) {               // pre-zero flood depth (pass B: free surface), final volume
            a.d_flood[i] = df;
            a.vol_flood[i] = noflow ? 0.0 : vf;
        }
        return;
    }
    double S = k.s64 ? *reinterpret_cast<const double *>(st + k.oS + e * 8)
                     : (double)*reinterpret_cast<const float *>(st + k.oS + e * 4);
    const double rough = *reinterpret_cast<const double *>(st + k.oRough + e * 8);
    double slope_t = S;                                              // kinematic: S_bed as it is
    if (VAR == EXT_DWB) {
        // ---- update_free_surface_slope :2583-2618: new depths of the cell and its D8 parent
        const double *dc = reinterpret_cast<const double *>(st + k.oQ) + (row + 1) * QW + ct + 2;
        d = dc[0];
        int dr, dcol;
        code_offset(code, dr, dcol);
        double d_par = dc[dr * QW + dcol];
        // a noflow cell's "parent" is cell (0, 0) of the GLOBAL grid (parent ID 0, :2590-2607):
        // its pre-zero depth where this slab can see that row, 0 otherwise (as the generic kernel)
        if (noflow) d_par = k.d00_visible ? __ldg(a.d + k.d00_index) : 0.0;
        double d_tot = d;
        if (k.flood) {               // free surface = channel + flood depth :2593-2596
            const double *fc = reinterpret_cast<const double *>(st + k.oDf) + (row + 1) * QW + ct + 2;
            df = fc[0];
            double df_par = fc[dr * QW + dcol];
            if (noflow) df_par = k.d00_visible ? __ldg(a.d_flood + k.d00_index) : 0.0;
            d_par = d_par + df_par;
            d_tot = d + df;
        }
        S = S + div_sel(d_tot - d_par, ds);                          // :2606-2607
        slope_t = fabs(S);
        if (r + a.row0 == a.outlet_row && c == a.outlet_col) qself = __ldg(a.Q_in + i);
    }
    // ---- update_trapezoid_Rh :2668-2703 ------------------------------------------------------
    const double A_wet = d * (w + (d * tc.x));
    double P_wet = w + div_sel(2.0 * d, tc.y);
    P_wet = noflow ? 1.0 : P_wet;
    double Rh = div_sel(A_wet, P_wet);
    Rh = noflow ? 0.0 : Rh;
    // ---- velocity: manning_formula :3960-4003 / law_of_the_wall :4005-4063 ---------------------
    double u, smooth = 0.0;
    if (!LOW) {
        u = div_sel(pow23_sel(Rh) * sqrt_sel(fabs(S)), rough);
    } else {
        smooth = np_max(div_sel(k.aval, rough) * d, 1.1);
        u = (k.law_const * sqrt_sel(Rh * fabs(S))) * log(smooth);
    }
    u = noflow ? 0.0 : u;                                            // :2862
    double f_cell = 0.0;
    const bool is_outlet = (r + a.row0 == a.outlet_row && c == a.outlet_col);
    if (DIAG || is_outlet) {
        if (d > 0.0) {                                               // :2762-2788
            if (!LOW) f_cell = k.g * ((rough * rough) / cbrt(d));
            else { const double qq = k.kappa / log(smooth); f_cell = qq * qq; }
        }
    }
    if (is_outlet) {                                                 // :2896-2904 (pre-zero depth)
        double *po = a.partials + kPartialStride;
        po[P_QOUT] = qself; po[P_UOUT] = u; po[P_DOUT] = d; po[P_FOUT] = f_cell; po[P_HASOUT] = 1.0;
    }
    if (DIAG) {                                                      // :2620-2641, :2867-2880
        const double tau = ((k.rho * k.g) * d) * slope_t;
        a.Rh[i] = Rh; a.A_wet[i] = A_wet; a.P_wet[i] = P_wet; a.f[i] = f_cell; a.tau[i] = tau;
        a.u_star[i] = sqrt(tau / k.rho);
        a.froude[i] = (d > 0.0) ? (u / sqrt(k.g * d)) : 0.0;
        if (VAR == EXT_DWB && a.S_free) a.S_free[i] = S;
    }
    double Qn = u * A_wet;                                           // next step :1697
    if (k.flood) {
        // update_flood_discharge :1701-1750 + update_discharge :1791-1862 for the NEXT step: the
        // flood depth after the noflow zeroing (:2968-2970), S_bed or this step's S_free
        const double dfz = noflow ? 0.0 : df;
        const double uf = div_sel(pow23_sel(dfz) * sqrt_sel(fabs(S)), k.flood_n);
        const double Af = (k.width_ratio * w) * dfz;
        Qn = Qn + (uf * Af);
        if (VAR == EXT_KW) { a.d_flood[i] = dfz; a.vol_flood[i] = noflow ? 0.0 : vf; }
    }
    if (VAR == EXT_KW && noflow) { acc.s[S_VOL_EDGE] += v; v = 0.0; d = 0.0; }   // :2960-2966
    const double dz = noflow ? 0.0 : d;
    if (!(dz >= 0.0 && dz <= 1.7976931348623157e308 && u >= 0.0 && u <= 1.7976931348623157e308)) {
        if (dz < 0.0) acc.n[S_NNEG_D - kNSum] += 1u;                 // :3193-3279, :3352-3454
        if (dz != dz) acc.n[S_NNAN_D - kNSum] += 1u;
        if (isinf(dz)) acc.n[S_NINF_D - kNSum] += 1u;
        if (u < 0.0) acc.n[S_NNEG_U - kNSum] += 1u;
        if (u != u) acc.n[S_NNAN_U - kNSum] += 1u;
        if (isinf(u)) acc.n[S_NINF_U - kNSum] += 1u;
    }
    rate_max = fmaxf(rate_max, __fdividef((float)u, (float)ds));     // CFL diagnostic: 1 / dt_min
    if (VAR == EXT_KW) { a.vol_out[i] = v; a.d[i] = d; }
    a.u[i] = u;
    a.Q_out[i] = Qn;
}

template <int VAR, bool LOW, bool DIAG, class TC>
__global__ void __launch_bounds__(TC::threads, 1)
channels_ext_kernel(const tfb_channels_step_args a, const tfb_channels_packed p, const ExtConsts k,
                    const __grid_constant__ ExtMaps maps) {
    constexpr int kTR = TC::R, kTW = TC::W, kQW = TC::QW, kConsumerWarps = TC::warps;
    constexpr int kMaxStages = 6;
    static_assert(TC::cpl == 1, "one cell per lane");
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ double sm[kConsumerWarps][kPartialStride];
    __shared__ double tot[kPartialStride];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
    __shared__ bool is_last;
    __shared__ volatile int s_fail;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NS = k.n_stages;
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
        if (lane == 0) {                              // producer: TMA loads of every stream
            int stage = 0;
            unsigned phase = 0;
            const int i0 = k.reverse ? (k.n_tiles - 1 - (int)blockIdx.x) : (int)blockIdx.x;
            int ty = i0 / k.tiles_x, txi = i0 - ty * k.tiles_x;
            for (int t = blockIdx.x; t < k.n_tiles; t += gridDim.x) {
                if (!mbar_wait(&empty_bar[stage], phase ^ 1u)) { s_fail = 1; break; }
                const int c0 = txi * kTW, r0 = ty * kTR;
                unsigned char *st = smem + (size_t)stage * k.stage_bytes;
                uint64_t *bar = &full_bar[stage];
                mbar_arrive_expect_tx(bar, k.tx_bytes);
                // state tensors are mapped from the top of their halo'd allocation
                tma_load_2d(st + k.oQ, &maps.Q, bar, c0 - 2, r0 - 1 + a.halo);
                tma_load_2d(st + k.oGeo, &maps.geo, bar, c0, r0);
                tma_load_2d(st + k.oW, &maps.w, bar, c0, r0);
                tma_load_2d(st + k.oRG, &maps.rg, bar, 0, r0);
                if (k.oSinu >= 0) tma_load_2d(st + k.oSinu, &maps.sinu, bar, c0, r0);
                if (VAR != EXT_DWB) {
                    tma_load_2d(st + k.oVol, &maps.vol, bar, c0, r0 + a.halo);
#pragma unroll
                    for (int j = 0; j < F_N; ++j)
                        if (k.foff[j] >= 0) tma_load_2d(st + k.foff[j], &maps.F[j], bar, c0, r0 + a.halo);
                    if (k.atten) tma_load_2d(st + k.oVs, &maps.vs, bar, c0, r0 + a.halo);
                    if (k.flood && k.oVb >= 0) tma_load_2d(st + k.oVb, &maps.vb, bar, c0, r0 + a.halo);
                } else if (k.flood) {
                    tma_load_2d(st + k.oDf, &maps.df, bar, c0 - 2, r0 - 1 + a.halo);
                }
                if (VAR != EXT_DWA) {
                    tma_load_2d(st + k.oS, &maps.S, bar, c0, r0);
                    tma_load_2d(st + k.oRough, &maps.rough, bar, c0, r0);
                }
                if (++stage == NS) { stage = 0; phase ^= 1u; }
                if (k.reverse) {                          // next tile of this CTA, no division
                    txi -= (int)gridDim.x;
                    while (txi < 0) { txi += k.tiles_x; --ty; }
                } else {
                    txi += (int)gridDim.x;
                    while (txi >= k.tiles_x) { txi -= k.tiles_x; ++ty; }
                }
            }
        }
    } else {
        int stage = 0;
        unsigned phase = 0;
        const int i0 = k.reverse ? (k.n_tiles - 1 - (int)blockIdx.x) : (int)blockIdx.x;
        int ty = i0 / k.tiles_x, txi = i0 - ty * k.tiles_x;
        for (int t = blockIdx.x; t < k.n_tiles; t += gridDim.x) {
            if (!mbar_wait(&full_bar[stage], phase)) { s_fail = 1; break; }
            const unsigned char *st = smem + (size_t)stage * k.stage_bytes;
#pragma unroll 1
            for (int rr = 0; rr < TC::rpw; ++rr) {
                const int row = warp / TC::wpr + rr * TC::RP, ct = ((warp % TC::wpr) << 5) + lane;
                const int c = txi * kTW + ct, r = ty * kTR + row;
                if (r < a.ny && c < a.nx)
                    ext_cell<VAR, LOW, DIAG>(a, p, k, st, kQW, kTW, row, ct, r, c, acc, rate_max);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);
            if (++stage == NS) { stage = 0; phase ^= 1u; }
            if (k.reverse) {
                txi -= (int)gridDim.x;
                while (txi < 0) { txi += k.tiles_x; --ty; }
            } else {
                txi += (int)gridDim.x;
                while (txi >= k.tiles_x) { txi -= k.tiles_x; ++ty; }
            }
        }
    }
    pk_finish<kConsumerWarps>(a, acc, rate_max, warp < kConsumerWarps, &s_fail,
                              VAR == EXT_KW ? 0 : (VAR == EXT_DWA ? 1 : 2), PkPeer{}, sm, tot, &is_last);
}

}  // namespace tfb

using namespace tfb;

namespace {

#ifndef TFB_EXT_ROWS
#define TFB_EXT_ROWS 15
#endif
#ifndef TFB_EXT_RPW
#define TFB_EXT_RPW 1
#endif
typedef TileCfg<TFB_EXT_ROWS, 1, 64, TFB_EXT_RPW> CfgExt;

template <int VAR, bool LOW, bool DIAG>
int launch_ext(const tfb_channels_step_args &a, const tfb_channels_packed &p, ExtConsts k,
               cudaStream_t st) {
    typedef CfgExt TC;
    constexpr int kTR = TC::R, kQR = TC::QR, kTW = TC::W, kQW = TC::QW;
    k.tiles_x = (a.nx + kTW - 1) / kTW;
    k.tiles_y = (a.ny + kTR - 1) / kTR;
    k.n_tiles = k.tiles_x * k.tiles_y;
    // kinematic: the caller alternates TFB_CH_REVERSE; diffusive: it sets the bit for pass B only
    k.reverse = (VAR != EXT_DWA && (a.flags & TFB_CH_REVERSE)) ? 1 : 0;
    // ---- ring slot laid out for the streams this step really has ---------------------------
    int off = 0;
    unsigned tx = 0;
    auto put = [&](int bytes_padded, int bytes_tma) { const int o = off; off += bytes_padded; tx += (unsigned)bytes_tma; return o; };
    k.oQ = put(TC::szQ, TC::nQ);
    k.oGeo = put(TC::szF, TC::nF);
    k.oW = k.w64 ? put(TC::szD, TC::nD) : put(TC::szF, TC::nF);
    k.oSinu = p.sinu64 ? put(TC::szD, TC::nD) : -1;
    k.oRG = put(((kTR * kRowGeom * 8 + 127) / 128) * 128, kTR * kRowGeom * 8);
    k.oVol = k.oS = k.oRough = 0;
    const tfb_sg *F[F_N] = {&a.P_rain, &a.SM, &a.GW, &a.MR, &a.ET, &a.IN};
    if (VAR != EXT_DWB) {
        k.oVol = put(TC::szD, TC::nD);
        for (int j = 0; j < F_N; ++j) {
            k.fval[j] = F[j]->val;
            k.foff[j] = F[j]->ptr ? put(TC::szD, TC::nD) : -1;
        }
        if (k.atten) k.oVs = put(TC::szD, TC::nD);
        k.oVb = (k.flood && a.vol_bankfull.ptr) ? put(TC::szD, TC::nD) : -1;
    } else if (k.flood) {
        k.oDf = put(TC::szQ, TC::nQ);
    }
    if (VAR != EXT_DWA) {
        k.oS = k.s64 ? put(TC::szD, TC::nD) : put(TC::szF, TC::nF);
        k.oRough = put(TC::szD, TC::nD);
    }
    k.stage_bytes = off;
    k.tx_bytes = tx;
    k.n_stages = kSmemBudget / off;
    if (k.n_stages > 6) k.n_stages = 6;
    if (k.n_stages < 2) { tfb::set_error("ext step: ring slot of %d bytes does not fit twice", off); return TFB_ERR_ARG; }
    // ---- tensor maps ---------------------------------------------------------------------
    ExtMaps m;
    memset(&m, 0, sizeof(m));
    const int64_t rows_h = (int64_t)a.ny + 2 * a.halo;
    const size_t hb = (size_t)a.halo * a.pitch;
    const CUtensorMapDataType F64 = CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    int rc;
    const double *halo_src = (VAR == EXT_DWB) ? a.d : a.Q_in;
    if ((rc = pk_get_map(halo_src - hb, rows_h, a.nx, a.pitch, F64, 8, kQW, kQR, &m.Q))) return rc;
    if ((rc = pk_get_map(p.geo32, a.ny, a.nx, a.pitch, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, kTW, kTR, &m.geo))) return rc;
    if (k.w64) {
        if ((rc = pk_get_map(p.width64, a.ny, a.nx, a.pitch, F64, 8, kTW, kTR, &m.w))) return rc;
    } else {
        if ((rc = pk_get_map(p.width32, a.ny, a.nx, a.pitch, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, kTW, kTR, &m.w))) return rc;
    }
    if (p.sinu64)
        if ((rc = pk_get_map(p.sinu64, a.ny, a.nx, a.pitch, F64, 8, kTW, kTR, &m.sinu))) return rc;
    if ((rc = pk_get_map(p.row_geom, a.ny, kRowGeom, kRowGeom, F64, 8, kRowGeom, kTR, &m.rg))) return rc;
    if (VAR != EXT_DWB) {
        if ((rc = pk_get_map(a.vol - hb, rows_h, a.nx, a.pitch, F64, 8, kTW, kTR, &m.vol))) return rc;
        for (int j = 0; j < F_N; ++j)
            if (F[j]->ptr)
                if ((rc = pk_get_map(F[j]->ptr - hb, rows_h, a.nx, a.pitch, F64, 8, kTW, kTR, &m.F[j]))) return rc;
        if (k.atten)
            if ((rc = pk_get_map(a.vol_stored - hb, rows_h, a.nx, a.pitch, F64, 8, kTW, kTR, &m.vs))) return rc;
        if (k.oVb >= 0)
            if ((rc = pk_get_map(a.vol_bankfull.ptr - hb, rows_h, a.nx, a.pitch, F64, 8, kTW, kTR, &m.vb))) return rc;
    } else if (k.flood) {
        if ((rc = pk_get_map(a.d_flood - hb, rows_h, a.nx, a.pitch, F64, 8, kQW, kQR, &m.df))) return rc;
    }
    if (VAR != EXT_DWA) {
        if (k.s64) {
            if ((rc = pk_get_map(p.S64, a.ny, a.nx, a.pitch, F64, 8, kTW, kTR, &m.S))) return rc;
        } else {
            if ((rc = pk_get_map(p.S32, a.ny, a.nx, a.pitch, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, kTW, kTR, &m.S))) return rc;
        }
        if ((rc = pk_get_map(p.rough64, a.ny, a.nx, a.pitch, F64, 8, kTW, kTR, &m.rough))) return rc;
    }
    const int smem = k.n_stages * k.stage_bytes;
    static bool attr_set[64] = {false};
    int dev = 0;
    TFB_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev >= 64 || !attr_set[dev]) {
        TFB_CUDA_CHECK(cudaFuncSetAttribute(channels_ext_kernel<VAR, LOW, DIAG, TC>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
        if (dev < 64) attr_set[dev] = true;
    }
    int64_t grid = sm_count();
    if (grid > k.n_tiles) grid = k.n_tiles;
    channels_ext_kernel<VAR, LOW, DIAG, TC><<<(int)grid, TC::threads, smem, st>>>(a, p, k, m);
    TFB_CUDA_CHECK(cudaGetLastError());
    return TFB_OK;
}

// argument checks shared by the three entry points
int prepare_ext(const tfb_channels_step_args *ap, tfb_channels_step_args &a,
                const tfb_channels_packed *pp, bool diffusive, ExtConsts &k) {
    TFB_ARG_CHECK(ap != nullptr && pp != nullptr);
    a = *ap;
    if (a.pitch == 0) a.pitch = a.nx;
    const tfb_channels_packed &p = *pp;
    TFB_ARG_CHECK(a.ny > 0 && a.nx > 0 && a.halo >= 0 && a.pitch >= a.nx && (a.pitch % 4) == 0);
    TFB_ARG_CHECK(p.geo32 && (p.width32 || p.width64) && p.lut_tc && p.row_geom &&
                  (p.S32 || p.S64) && p.rough64 && p.n_angles >= 1 && p.n_angles <= 65536);
    TFB_ARG_CHECK(a.vol && a.vol_out && a.Q_in && a.Q_out && a.d && a.u);
    TFB_ARG_CHECK(a.Q_in != a.Q_out && a.scalars && a.partials && a.ticket && a.dt > 0.0);
    TFB_ARG_CHECK(!(a.flags & (TFB_CH_SRC_MET | TFB_CH_SRC_SNOW | TFB_CH_SRC_EVAP | TFB_CH_SRC_INFIL)));
    if (a.flags & TFB_CH_FLOOD)
        TFB_ARG_CHECK(a.d_flood && a.vol_flood && a.width_ratio > 0.0 && a.flood_manning_n > 0.0);
    if (a.flags & TFB_CH_ATTENUATE)
        TFB_ARG_CHECK(a.vol_stored && a.vol_stored_out && a.t_drain > 0.0);
    TFB_ARG_CHECK(((a.flags & TFB_CH_DIFFUSIVE) != 0) == diffusive);
    TFB_ARG_CHECK(!a.sinu.ptr || p.sinu64);        // a sinuosity grid comes as a packed stream
    if (a.flags & TFB_CH_DIAG)
        TFB_ARG_CHECK(a.Rh && a.A_wet && a.P_wet && a.f && a.tau && a.u_star && a.froude);
    if (a.flags & TFB_CH_WRITE_R) TFB_ARG_CHECK(a.R != nullptr);
    auto al16 = [](const void *q) { return (((uintptr_t)q) & 15u) == 0; };
    TFB_ARG_CHECK(al16(a.vol) && al16(a.Q_in) && al16(a.d) && al16(p.rough64) && al16(p.lut_tc) &&
                  al16(p.geo32) && al16(p.width32) && al16(p.S32) && al16(p.width64) && al16(p.S64) &&
                  al16(p.sinu64) && al16(a.P_rain.ptr) &&
                  al16(a.SM.ptr) && al16(a.GW.ptr) && al16(a.MR.ptr) && al16(a.ET.ptr) && al16(a.IN.ptr) &&
                  al16(a.vol_stored) && al16(a.vol_bankfull.ptr) && al16(a.d_flood));
    memset(&k, 0, sizeof(k));
    k.g = 9.81; k.aval = 0.476; k.kappa = 0.408;
    k.law_const = sqrt(k.g) / k.kappa;                   // channels_base.py:528-531
    k.dt = a.dt;
    k.rho = a.rho_H2O;
    k.et_grid = a.ET.ptr ? 1 : 0;
    k.r_grid = a.is_R_grid ? 1 : 0;
    k.write_R = (a.flags & TFB_CH_WRITE_R) ? 1 : 0;
    if (!(a.flags & TFB_CH_ET_INPLACE)) a.ET_rw = nullptr;
    k.d00_visible = (-a.row0 >= -a.halo && -a.row0 < a.ny + a.halo) ? 1 : 0;
    k.d00_index = (long long)(-a.row0) * a.pitch;
    k.atten = (a.flags & TFB_CH_ATTENUATE) ? 1 : 0;
    k.flood = (a.flags & TFB_CH_FLOOD) ? 1 : 0;
    k.oVs = k.oDf = 0; k.oVb = -1;
    k.w64 = p.width32 ? 0 : 1;
    k.s64 = p.S32 ? 0 : 1;
    k.oSinu = -1;
    k.t_drain = a.t_drain; k.vb_val = a.vol_bankfull.val;
    k.width_ratio = a.width_ratio; k.flood_n = a.flood_manning_n;
    const bool any_grid = a.P_rain.ptr || a.SM.ptr || a.GW.ptr || a.MR.ptr || a.ET.ptr || a.IN.ptr;
    TFB_ARG_CHECK(!any_grid || a.is_R_grid == 1);
    return TFB_OK;
}

template <int VAR>
int dispatch_ext(const tfb_channels_step_args &a, const tfb_channels_packed &p, const ExtConsts &k,
                 cudaStream_t st) {
    const bool low = (a.flags & TFB_CH_LAW_OF_WALL) != 0, diag = (a.flags & TFB_CH_DIAG) != 0;
    if constexpr (VAR == EXT_DWA) {
        return launch_ext<EXT_DWA, false, false>(a, p, k, st);     // pass A has no velocity law
    } else {
        if (low && diag) return launch_ext<VAR, true, true>(a, p, k, st);
        if (low) return launch_ext<VAR, true, false>(a, p, k, st);
        if (diag) return launch_ext<VAR, false, true>(a, p, k, st);
        return launch_ext<VAR, false, false>(a, p, k, st);
    }
}

}  // namespace

extern "C" int tfb_channels_step_ext(const tfb_channels_step_args *ap, const tfb_channels_packed *pp,
                                     void *stream) {
    tfb_channels_step_args a;
    ExtConsts k;
    const int rc = prepare_ext(ap, a, pp, false, k);
    if (rc) return rc;
    return dispatch_ext<EXT_KW>(a, *pp, k, (cudaStream_t)stream);
}

extern "C" int tfb_channels_step_ext_dw_a(const tfb_channels_step_args *ap,
                                          const tfb_channels_packed *pp, void *stream) {
    tfb_channels_step_args a;
    ExtConsts k;
    const int rc = prepare_ext(ap, a, pp, true, k);
    if (rc) return rc;
    TFB_ARG_CHECK(a.vol != a.vol_out);
    return dispatch_ext<EXT_DWA>(a, *pp, k, (cudaStream_t)stream);
}

extern "C" int tfb_channels_step_ext_dw_b(const tfb_channels_step_args *ap,
                                          const tfb_channels_packed *pp, void *stream) {
    tfb_channels_step_args a;
    ExtConsts k;
    const int rc = prepare_ext(ap, a, pp, true, k);
    if (rc) return rc;
    return dispatch_ext<EXT_DWB>(a, *pp, k, (cudaStream_t)stream);
}
