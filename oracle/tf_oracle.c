/*
 * tf_oracle.c -- plain-C restatement of the TopoFlow 3.6 hot path (CPU).
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_build/
 * libtf_oracle.so and loaded with ctypes by oracle/c_oracle.py.  Only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may use it; the
 * product (topoflow36_b200) never links or calls it.
 *
 * It exists because the numpy oracle (oracle/np_oracle.py, bit-exact against
 * the reference) needs ~45 fp64 temporaries per grid and ~0.1 s per Mcell-step;
 * this file does the same arithmetic in the same order with two passes over
 * the grid, optionally on all host cores (pthreads over row chunks), so that 4096^2 parity runs
 * and the CPU baseline finish in seconds.  Compile with -ffp-contract=off: no
 * fused multiply-add, so every + - * / sqrt rounds exactly as numpy's.  pow()
 * and log() are glibc's (<1 ulp), numpy's are SVML/libm (<= 1 ulp): results
 * agree with np_oracle to ~1e-15 relative, checked in tests/test_c_oracle.py.
 *
 * Parity status: PINNED through np_oracle (which is bit-exact against the
 * reference fixtures) -- see tests/test_c_oracle.py.
 *
 * Citations: components/channels_base.py (file:line given per block),
 * components/d8_global.py, components/infil_green_ampt.py, snow_degree_day.py,
 * evap_priestley_taylor.py, met_base.py of /root/reference/topoflow.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { const double *ptr; double val; } tfo_sg;

#define SG(s, i) ((s).ptr ? (s).ptr[(i)] : (s).val)

static inline double np_max(double a, double b) { return (a >= b || a != a) ? a : b; }
static inline double np_min(double a, double b) { return (a <= b || a != a) ? a : b; }

/* D8: k = 0..7 <-> NE,E,SE,S,SW,W,NW,N  (d8_base.py:561-640) */
static const int NBR_DR[8] = {-1, 0, 1, 1, 1, 0, -1, -1};
static const int NBR_DC[8] = {1, 1, 1, 0, -1, -1, -1, 0};

static inline float ds_of(unsigned code, float dx, float dy, float dd) {
    /* d8_global.update_flow_length_grid :960-1027 */
    if (code == 1 || code == 4 || code == 16 || code == 64) return dd;
    if (code == 8 || code == 128) return dy;
    return dx;                       /* E, W and code 0 */
}

static inline void code_offset(unsigned code, int *dr, int *dc) {
    *dr = 0; *dc = 0;
    for (int k = 0; k < 8; ++k)
        if (code == (1u << k)) { *dr = NBR_DR[k]; *dc = NBR_DC[k]; }
}

typedef struct {
    int32_t ny, nx;
    int32_t diffusive, law_of_wall, fused;      /* fused: bit0 met, 1 snow, 2 evap, 3 infil */
    int32_t outlet_row, outlet_col;
    int32_t is_R_grid;
    double dt, time_min, rho_H2O;
    const uint8_t *code;
    const float *dx, *dy, *dd;                  /* per row */
    tfo_sg sinu, da_row, width, angle, rough;   /* da_row.ptr = per-ROW array */
    const double *S_bed;
    double *vol, *Q_in, *Q_out, *d, *u;         /* vol in/out */
    double *d_pre;                              /* scratch: depth before noflow zeroing */
    double *Rh, *A_wet, *P_wet, *f, *tau, *u_star, *froude, *S_free, *R;  /* optional */
    tfo_sg P_rain, SM, GW, MR, IN, ET; double *ET_rw;
    tfo_sg T_air; double T_rain_snow;
    tfo_sg c0, T0; double *h_swe, *h_snow; double rho_snow;
    tfo_sg T_surf, Qn_SW, Qn_LW; double alpha, K_soil, soil_x, T_soil_x;
    tfo_sg Ks, Ki, qs, qi, G; double *I;
    /* scalar block, same meaning as tfb_channels_scalars */
    double vol_R, vol_Q, vol_edge, Q_outlet, u_outlet, d_outlet, f_outlet;
    double Q_peak, T_peak, u_peak, Tu_peak, d_peak, Td_peak;
    double vol_P, vol_PR, vol_PS, vol_SM, vol_ET, vol_IN;
    int64_t n_neg_d, n_nan_d, n_inf_d, n_neg_u, n_nan_u, n_inf_u;
} tfo_channels;

int64_t tfo_sizeof_channels(void) { return (int64_t)sizeof(tfo_channels); }

/* per-thread accumulators of one pass */
typedef struct {
    double sR, sP, sPR, sPS, sSM, sET, sIN, sEdge;
    int64_t nneg_d, nnan_d, ninf_d, nneg_u, nnan_u, ninf_u;
    double out[4]; int has_out;
} tfo_acc;

typedef struct { tfo_channels *c; int r0, r1; tfo_acc acc; int pass; } tfo_job;

static void pass1_rows(tfo_channels *c, int r0, int r1, tfo_acc *acc);
static void pass2_rows(tfo_channels *c, int r0, int r1, tfo_acc *acc);

static void *job_main(void *p) {
    tfo_job *j = (tfo_job *)p;
    if (j->pass == 1) pass1_rows(j->c, j->r0, j->r1, &j->acc);
    else pass2_rows(j->c, j->r0, j->r1, &j->acc);
    return NULL;
}

/* run one pass over row chunks on `nthreads` pthreads; partial sums are folded
 * in chunk order, so a given thread count always gives the same result */
static void run_pass(tfo_channels *c, int pass, int nthreads, tfo_acc *total) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > c->ny) nthreads = c->ny;
    tfo_job *jobs = (tfo_job *)calloc((size_t)nthreads, sizeof(tfo_job));
    pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; ++t) {
        jobs[t].c = c; jobs[t].pass = pass;
        jobs[t].r0 = (int)((int64_t)c->ny * t / nthreads);
        jobs[t].r1 = (int)((int64_t)c->ny * (t + 1) / nthreads);
        if (t > 0) pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    job_main(&jobs[0]);
    for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
    memset(total, 0, sizeof(*total));
    for (int t = 0; t < nthreads; ++t) {
        const tfo_acc *a = &jobs[t].acc;
        total->sR += a->sR; total->sP += a->sP; total->sPR += a->sPR; total->sPS += a->sPS;
        total->sSM += a->sSM; total->sET += a->sET; total->sIN += a->sIN; total->sEdge += a->sEdge;
        total->nneg_d += a->nneg_d; total->nnan_d += a->nnan_d; total->ninf_d += a->ninf_d;
        total->nneg_u += a->nneg_u; total->nnan_u += a->nnan_u; total->ninf_u += a->ninf_u;
        if (a->has_out) { memcpy(total->out, a->out, sizeof(a->out)); total->has_out = 1; }
    }
    free(jobs); free(th);
}

/* ---- pass 1: sources, R, new volume, new depth ------------------------- */
static void pass1_rows(tfo_channels *c, int r0, int r1, tfo_acc *acc)
{
    const int ny = c->ny, nx = c->nx;
    const double dt = c->dt;
    double sR = 0, sP = 0, sPR = 0, sPS = 0, sSM = 0, sET = 0, sIN = 0;
    for (int r = r0; r < r1; ++r) {
        const double da = c->da_row.ptr ? c->da_row.ptr[r] : c->da_row.val;
        for (int col = 0; col < nx; ++col) {
            const int64_t i = (int64_t)r * nx + col;
            double P_rain = SG(c->P_rain, i), SMv = SG(c->SM, i), INv = SG(c->IN, i);
            double ETv = SG(c->ET, i);
            int ET_is_grid = (c->ET.ptr != NULL);
            double P_snow = 0.0;
            if (c->fused & 1) {                          /* met_base.py:1300-1388 */
                const double P = SG(c->P_rain, i), Ta = SG(c->T_air, i);
                P_rain = P * ((Ta > c->T_rain_snow) ? 1.0 : 0.0);
                P_snow = P * ((Ta <= c->T_rain_snow) ? 1.0 : 0.0);
                sP += (P * da) * dt; sPR += (P_rain * da) * dt; sPS += (P_snow * da) * dt;
            }
            if (c->fused & 2) {                          /* snow_degree_day.py:293-360 */
                const double h = c->h_swe[i];
                double M = (SG(c->c0, i) / 8.64E7) * (SG(c->T_air, i) - SG(c->T0, i));
                M = np_min(M, h / dt);                   /* snow_base.py:490-512 */
                M = np_max(M, 0.0);
                SMv = M;
                double hn = h + (P_snow * dt);           /* snow_base.py:559-609 */
                hn = hn - (M * dt);
                hn = np_max(hn, 0.0);
                c->h_swe[i] = hn;
                if (c->h_snow) c->h_snow[i] = hn * (c->rho_H2O / c->rho_snow);
                sSM += (M * da) * dt;
            }
            if (c->fused & 4) {                          /* evap_priestley_taylor.py:308-413 */
                const double Ta = SG(c->T_air, i);
                const double Qc = c->K_soil * (SG(c->T_surf, i) - c->T_soil_x) / c->soil_x;
                const double Qnet = SG(c->Qn_SW, i) + SG(c->Qn_LW, i);
                const double Qet = (c->alpha * (0.406 + (0.011 * Ta))) * (Qnet + Qc);
                ETv = np_max(Qet / 2.5E+9, 0.0);
                ET_is_grid = 1;
                sET += (ETv * da) * dt;
            }
            if (c->fused & 8) {                          /* infil_green_ampt.py:511-578 */
                const double Ks = SG(c->Ks, i), P_total = P_rain + SMv;
                const double dK = Ks - SG(c->Ki, i), dq = SG(c->qs, i) - SG(c->qi, i);
                const double Iv = c->I[i];
                const double fc = (((SG(c->G, i) * dK) * dq) / Iv) + Ks;
                double in = np_min(fc, P_total);
                if (P_total < Ks) in = P_total;          /* infil_base.py:954-995 */
                INv = in;
                c->I[i] = Iv + (in * dt);                /* infil_base.py:797 */
                sIN += (in * da) * dt;
            }
            /* update_R :1625-1649 */
            const double vol_old = c->vol[i];
            if (ET_is_grid) {
                if (vol_old < ((ETv * da) * dt)) {
                    ETv = 0.0;
                    if (c->ET_rw && !(c->fused & 4)) c->ET_rw[i] = 0.0;
                }
            } else {
                ETv = 0.0;
            }
            const double R = (((P_rain + SMv) + SG(c->GW, i)) + SG(c->MR, i)) - (ETv + INv);
            if (c->R) c->R[i] = R;
            if (c->is_R_grid) sR += (R * da) * dt;       /* :1666-1670 */
            /* update_flow_volume :2086-2145 -- parent-side gather, children in
               the reference's scatter order (codes 1,2,4,...,128) */
            double v = vol_old + ((R * da) * dt);
            for (int k = 0; k < 8; ++k) {
                /* a child with code 1<<k sits opposite to direction k */
                const int rr = r - NBR_DR[k], cc = col - NBR_DC[k];
                if (rr < 0 || rr >= ny || cc < 0 || cc >= nx) continue;
                const int64_t j = (int64_t)rr * nx + cc;
                if (c->code[j] == (uint8_t)(1u << k)) v += dt * c->Q_in[j];
            }
            v -= (c->Q_in[i] * dt);
            v = np_max(v, 0.0);
            c->vol[i] = v;   /* safe in place: nobody else reads vol in pass 1 */
            /* update_channel_depth :2299-2330, :2379 */
            const double ang = SG(c->angle, i), w = SG(c->width, i);
            const double ds = SG(c->sinu, i) *
                (double)ds_of(c->code[i], c->dx[r], c->dy[r], c->dd[r]);   /* :1242 */
            double d;
            if (ang == 0.0) {
                d = v / (w * ds);
            } else {
                const double denom = 2.0 * tan(ang);
                double arg = ((2.0 * denom) * v) / ds;
                arg += w * w;
                d = (sqrt(arg) - w) / denom;
            }
            c->d_pre[i] = np_max(d, 0.0);
        }
    }
    acc->sR = sR; acc->sP = sP; acc->sPR = sPR; acc->sPS = sPS;
    acc->sSM = sSM; acc->sET = sET; acc->sIN = sIN;
}

/* ---- pass 2: Rh, slope, velocity, diagnostics, edges, checks ------------ */
static void pass2_rows(tfo_channels *c, int r0, int r1, tfo_acc *acc)
{
    const int nx = c->nx;
    const double g = 9.81, aval = 0.476, kappa = 0.408;
    const double law_const = sqrt(g) / kappa;            /* :531 */
    const double one_third = 1.0 / 3.0, two_thirds = 2.0 / 3.0;
    double sEdge = 0;
    int64_t nneg_d = 0, nnan_d = 0, ninf_d = 0, nneg_u = 0, nnan_u = 0, ninf_u = 0;
    const double d00 = c->d_pre[0];
    for (int r = r0; r < r1; ++r) {
        for (int col = 0; col < nx; ++col) {
            const int64_t i = (int64_t)r * nx + col;
            const unsigned code = c->code[i];
            const int noflow = (code == 0);
            double d = c->d_pre[i];
            const double ang = SG(c->angle, i), w = SG(c->width, i);
            const double rough = SG(c->rough, i);
            /* update_trapezoid_Rh :2668-2703 */
            const double L2 = d * tan(ang);
            const double A_wet = d * (w + L2);
            double P_wet = w + ((2.0 * d) / cos(ang));
            if (noflow) P_wet = 1.0;
            double Rh = A_wet / P_wet;
            if (noflow) Rh = 0.0;
            const double S_bed = c->S_bed[i];
            double S = S_bed;
            if (c->diffusive) {                          /* :2606-2607 */
                const double ds = SG(c->sinu, i) *
                    (double)ds_of(code, c->dx[r], c->dy[r], c->dd[r]);
                double dpar = d00;                       /* noflow -> parent (0,0) */
                if (!noflow) {
                    int dr, dc; code_offset(code, &dr, &dc);
                    dpar = c->d_pre[(int64_t)(r + dr) * nx + (col + dc)];
                }
                S = S_bed + ((d - dpar) / ds);
                if (c->S_free) c->S_free[i] = S;
            }
            const double S2 = fabs(S);
            double u, smooth = 0.0;
            if (!c->law_of_wall) {                       /* :3982-3983 */
                u = (pow(Rh, two_thirds) * sqrt(S2)) / rough;
            } else {                                     /* :4038-4047 */
                smooth = np_max((aval / rough) * d, 1.1);
                u = (law_const * sqrt(Rh * S2)) * log(smooth);
            }
            if (noflow) u = 0.0;                         /* :2862 */
            double f = 0.0, fr = 0.0;
            if (d > 0.0) {                               /* :2762-2788, :2879 */
                if (!c->law_of_wall) f = g * ((rough * rough) / pow(d, one_third));
                else { const double q = kappa / log(smooth); f = q * q; }
                fr = u / sqrt(g * d);
            }
            if (c->Rh) {
                const double slope_t = c->diffusive ? S2 : S_bed;      /* :2626-2630 */
                const double tau = ((c->rho_H2O * g) * d) * slope_t;
                c->Rh[i] = Rh; c->A_wet[i] = A_wet; c->P_wet[i] = P_wet;
                c->f[i] = f; c->tau[i] = tau; c->u_star[i] = sqrt(tau / c->rho_H2O);
                c->froude[i] = fr;
            }
            if (r == c->outlet_row && col == c->outlet_col) {          /* :2896-2904 */
                acc->out[0] = c->Q_in[i]; acc->out[1] = u; acc->out[2] = d; acc->out[3] = f;
                acc->has_out = 1;
            }
            if (noflow) {                                /* update_edge_values :2960-2966 */
                sEdge += c->vol[i];
                c->vol[i] = 0.0;
                d = 0.0;
            }
            if (d < 0.0) ++nneg_d;
            if (d != d) ++nnan_d;
            if (isinf(d)) ++ninf_d;
            if (u < 0.0) ++nneg_u;
            if (u != u) ++nnan_u;
            if (isinf(u)) ++ninf_u;
            c->d[i] = d;
            c->u[i] = u;
            c->Q_out[i] = u * A_wet;                     /* next step's :1697 */
        }
    }
    acc->sEdge = sEdge;
    acc->nneg_d = nneg_d; acc->nnan_d = nnan_d; acc->ninf_d = ninf_d;
    acc->nneg_u = nneg_u; acc->nnan_u = nnan_u; acc->ninf_u = ninf_u;
}

/* One channels_component.update() (channels_base.py:803-962), non-flood path. */
int tfo_channels_step(tfo_channels *c, int nthreads)
{
    const int ny = c->ny, nx = c->nx;
    const double dt = c->dt;
    tfo_acc a1, a2;
    run_pass(c, 1, nthreads, &a1);
    run_pass(c, 2, nthreads, &a2);
    if (a2.has_out) {
        c->Q_outlet = a2.out[0]; c->u_outlet = a2.out[1];
        c->d_outlet = a2.out[2]; c->f_outlet = a2.out[3];
    }
    if (c->is_R_grid) {
        c->vol_R += a1.sR;
    } else {                                             /* :1666-1668 */
        const double R = (((c->P_rain.val + c->SM.val) + c->GW.val) + c->MR.val) - (0.0 + c->IN.val);
        c->vol_R += ((R * c->da_row.val) * dt) * (double)((int64_t)ny * nx);
    }
    c->vol_edge += a2.sEdge;
    c->vol_P += a1.sP; c->vol_PR += a1.sPR; c->vol_PS += a1.sPS;
    c->vol_SM += a1.sSM; c->vol_ET += a1.sET; c->vol_IN += a1.sIN;
    c->n_neg_d = a2.nneg_d; c->n_nan_d = a2.nnan_d; c->n_inf_d = a2.ninf_d;
    c->n_neg_u = a2.nneg_u; c->n_nan_u = a2.nnan_u; c->n_inf_u = a2.ninf_u;
    /* :2908-2934 */
    if (c->Q_outlet > c->Q_peak) { c->Q_peak = c->Q_outlet; c->T_peak = c->time_min; }
    if (c->u_outlet > c->u_peak) { c->u_peak = c->u_outlet; c->Tu_peak = c->time_min; }
    if (c->d_outlet > c->d_peak) { c->d_peak = c->d_outlet; c->Td_peak = c->time_min; }
    c->vol_Q += (c->Q_outlet * dt);
    return 0;
}

/* ------------------------------------------------------------------------
 * D8: start codes + tie break + flat linking + finalize
 * (d8_global.py:181-353, :355-372, :374-576, :139-146)
 * ---------------------------------------------------------------------- */
static inline float np_maxf(float a, float b) { return (a >= b || a != a) ? a : b; }

int tfo_d8_flow_grid(const float *dem, int ny, int nx, float dx, float dy, float dd,
                     double nodata, int link_flats, const uint8_t *resolve,
                     uint8_t *code_out, int *n_reps_out)
{
    const int64_t n = (int64_t)ny * nx;
    int16_t *code = (int16_t *)calloc((size_t)n, sizeof(int16_t));
    int16_t *next = (int16_t *)malloc((size_t)n * sizeof(int16_t));
    if (!code || !next) { free(code); free(next); return -1; }
    for (int r = 1; r < ny - 1; ++r) {
        for (int c = 1; c < nx - 1; ++c) {
            const int64_t i = (int64_t)r * nx + c;
            const float z = dem[i];
            float s[8];
            for (int k = 0; k < 8; ++k) {
                const float zn = dem[i + (int64_t)NBR_DR[k] * nx + NBR_DC[k]];
                const float dist = (k & 1) ? ((k == 1 || k == 5) ? dx : dy) : dd;
                s[k] = (z - zn) / dist;
            }
            float mx = np_maxf(s[0], s[1]);
            for (int k = 2; k < 8; ++k) mx = np_maxf(mx, s[k]);
            int cd = 0;
            for (int k = 0; k < 8; ++k) cd += (s[k] == mx) ? (1 << k) : 0;
            if (!link_flats) { if (mx <= 0.0f) cd = 0; }
            else { if (mx == 0.0f) cd = -cd; if (mx < 0.0f) cd = -300; }
            if ((double)z <= nodata || !isfinite(z)) cd = 0;
            if (cd > 0) cd = resolve[cd];                 /* break_flow_grid_ties */
            code[i] = (int16_t)cd;
        }
    }
    int reps = 0;
    if (link_flats) {
        while (1) {
            ++reps;
            int64_t n_flats = 0, n_fixed = 0;
            memcpy(next, code, (size_t)n * sizeof(int16_t));
            for (int r = 1; r < ny - 1; ++r) {
                for (int c = 1; c < nx - 1; ++c) {
                    const int64_t i = (int64_t)r * nx + c;
                    const int v = code[i];
                    if (!(v < 0 && v != -300)) continue;
                    ++n_flats;
                    const int tied = -v;
                    int ready = 0;
                    for (int k = 0; k < 8; ++k) {
                        const int nb = code[i + (int64_t)NBR_DR[k] * nx + NBR_DC[k]];
                        const int opp = 1 << ((k + 4) & 7);
                        if (nb > 0 && nb != opp && (tied & (1 << k))) ready |= (1 << k);
                    }
                    if (ready > 0) { next[i] = resolve[ready]; ++n_fixed; }
                }
            }
            if (n_flats == 0) break;
            int16_t *t = code; code = next; next = t;
            if (n_fixed == n_flats || n_fixed == 0) break;
        }
    }
    for (int64_t i = 0; i < n; ++i) {
        const int v = code[i];
        code_out[i] = (v > 0 && v <= 128 && (v & (v - 1)) == 0) ? (uint8_t)v : 0;
    }
    free(code); free(next);
    if (n_reps_out) *n_reps_out = reps;
    return 0;
}

/* Contributing area, O(N) Kahn order with the reference's per-cell fp32 sum in
 * neighbour order k = 0..7 (d8_global.py:1223-1258); optional exact counts. */
int tfo_d8_area(const uint8_t *code, int ny, int nx, double pixel_area,
                const double *pixel_area_row, float *A, int64_t *count)
{
    const int64_t n = (int64_t)ny * nx;
    int8_t *pending = (int8_t *)malloc((size_t)n);
    int64_t *queue = (int64_t *)malloc((size_t)n * sizeof(int64_t));
    if (!pending || !queue) { free(pending); free(queue); return -1; }
    int64_t head = 0, tail = 0;
    for (int r = 0; r < ny; ++r) for (int c = 0; c < nx; ++c) {
        const int64_t i = (int64_t)r * nx + c;
        A[i] = 0.0f;
        if (count) count[i] = 0;
        int m = 0;
        if (code[i] != 0)
            for (int k = 0; k < 8; ++k) {
                const int rr = r + NBR_DR[k], cc = c + NBR_DC[k];
                if (rr < 0 || rr >= ny || cc < 0 || cc >= nx) continue;
                if (code[(int64_t)rr * nx + cc] == (uint8_t)(1u << ((k + 4) & 7))) ++m;
            }
        pending[i] = (int8_t)m;
        if (code[i] != 0 && m == 0) queue[tail++] = i;
    }
    while (head < tail) {
        const int64_t i = queue[head++];
        const int r = (int)(i / nx), c = (int)(i - (int64_t)r * nx);
        float a = (float)(pixel_area_row ? pixel_area_row[r] : pixel_area);
        int64_t cnt = 1;
        for (int k = 0; k < 8; ++k) {
            const int rr = r + NBR_DR[k], cc = c + NBR_DC[k];
            if (rr < 0 || rr >= ny || cc < 0 || cc >= nx) continue;
            const int64_t j = (int64_t)rr * nx + cc;
            if (code[j] == (uint8_t)(1u << ((k + 4) & 7))) { a += A[j]; if (count) cnt += count[j]; }
        }
        A[i] = a;
        if (count) count[i] = cnt;
        int dr, dc; code_offset(code[i], &dr, &dc);
        const int pr = r + dr, pc = c + dc;
        if (pr < 0 || pr >= ny || pc < 0 || pc >= nx) continue;
        const int64_t p = (int64_t)pr * nx + pc;
        if (code[p] == 0) continue;
        if (--pending[p] == 0) queue[tail++] = p;
    }
    free(pending); free(queue);
    return 0;
}

/* ---------------------------------------------------------------------------
 * Depression filling: utils/fill_pits.py fill_pits() :358-640, get_start_pixels()
 * :178-356 -- priority flood with a binary heap of (elevation, ID) pairs, ordered
 * like Python tuples.  close_nodata = 0 reproduces what the reference does with the
 * "closed" array (:404-409: the ROW numbers of the nodata cells are used as flat IDs);
 * see oracle/np_oracle.fill_pits.  In place; returns the number of raised cells.
 * ------------------------------------------------------------------------- */
typedef struct { float z; int64_t id; } tfo_hp;
static inline int hp_less(tfo_hp a, tfo_hp b) { return a.z < b.z || (a.z == b.z && a.id < b.id); }
static void hp_push(tfo_hp *h, int64_t *n, tfo_hp v) {
    int64_t i = (*n)++;
    while (i > 0) {
        const int64_t p = (i - 1) >> 1;
        if (!hp_less(v, h[p])) break;
        h[i] = h[p]; i = p;
    }
    h[i] = v;
}
static tfo_hp hp_pop(tfo_hp *h, int64_t *n) {
    const tfo_hp top = h[0], v = h[--(*n)];
    int64_t i = 0;
    for (;;) {
        int64_t c = 2 * i + 1;
        if (c >= *n) break;
        if (c + 1 < *n && hp_less(h[c + 1], h[c])) ++c;
        if (!hp_less(h[c], v)) break;
        h[i] = h[c]; i = c;
    }
    if (*n > 0) h[i] = v;
    return top;
}

int64_t tfo_fill_pits(float *dem, int ny, int nx, float nodata, float closed_code, int close_nodata)
{
    const int64_t n = (int64_t)ny * nx;
    uint8_t *C = (uint8_t *)calloc((size_t)n, 1);       /* 0 open, 1 closed, 2 on heap */
    uint8_t *seed = (uint8_t *)calloc((size_t)n, 1);
    tfo_hp *heap = (tfo_hp *)malloc((size_t)n * sizeof(tfo_hp));
    if (!C || !seed || !heap) { free(C); free(seed); free(heap); return -1; }
    for (int64_t i = 0; i < n; ++i) if (!isfinite(dem[i])) dem[i] = nodata;       /* :232-235 */
    for (int r = 0; r < ny; ++r)
        for (int c = 0; c < nx; ++c) {
            const int64_t i = (int64_t)r * nx + c;
            const float z = dem[i];
            if (r == 0 || r == ny - 1 || c == 0 || c == nx - 1) { seed[i] = z > nodata; continue; }
            const int my = ny - 2, mx = nx - 2;           /* rolls of the interior sub-array */
            float zs = INFINITY;
            for (int k = 0; k < 8; ++k) {
                int rr = r - 1 + NBR_DR[k], cc = c - 1 + NBR_DC[k];
                rr = (rr < 0) ? rr + my : ((rr >= my) ? rr - my : rr);
                cc = (cc < 0) ? cc + mx : ((cc >= mx) ? cc - mx : cc);
                const float v = dem[(int64_t)(rr + 1) * nx + cc + 1];
                if (v < zs) zs = v;
            }
            seed[i] = (z == closed_code) || (z > nodata && zs <= nodata);
        }
    for (int r = 0; r < ny; ++r)
        for (int c = 0; c < nx; ++c)
            if (dem[(int64_t)r * nx + c] <= nodata) {
                if (close_nodata) C[(int64_t)r * nx + c] = 1;
                else if (r < n) C[r] = 1;                 /* row number used as a flat ID */
            }
    int64_t nh = 0, n_raised = 0;
    for (int64_t i = 0; i < n; ++i)
        if (seed[i]) { C[i] = 2; tfo_hp v = {dem[i], i}; hp_push(heap, &nh, v); }
    const int64_t incs[8] = {-nx - 1, -nx, -nx + 1, -1, 1, nx - 1, nx, nx + 1};
    while (nh > 0) {
        const tfo_hp t = hp_pop(heap, &nh);
        C[t.id] = 1;
        for (int k = 0; k < 8; ++k) {
            const int64_t j = t.id + incs[k];             /* range check only (:523-541) */
            if (j < 0 || j >= n || C[j] != 0) continue;
            if (dem[j] < t.z) { dem[j] = t.z; ++n_raised; }
            C[j] = 2;
            tfo_hp v = {dem[j], j};
            hp_push(heap, &nh, v);
        }
    }
    free(C); free(seed); free(heap);
    return n_raised;
}
