/*
 * tfb.h -- C ABI of libtfb_b200.so: B200 (sm_100a) kernels for the TopoFlow 3.6
 * per-timestep grid update (channel routing + D8 toolkit + source terms).
 *
 * The reference (peckhams/topoflow36) is pure Python/numpy and has no FFI for
 * this path; each entry point below cites the reference method(s) it replaces
 * (paths relative to topoflow/).  The Python BMI components in
 * topoflow36_b200/ bind these symbols with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, a negative tfb_status otherwise and
 *     never throws; tfb_last_error() gives the text of the last failure on the
 *     calling thread;
 *   - all grid pointers are DEVICE pointers to row-major (ny, nx) arrays, row 0
 *     = north, caller-owned; `stream` is a cudaStream_t passed as void*;
 *   - "scalar-or-grid" inputs follow the reference's CFG convention
 *     (<var>_type = Scalar | Grid): a tfb_sg with ptr == NULL means the scalar
 *     `val` applies to every cell;
 *   - no global state except the opaque plan objects; calls on one plan must be
 *     serialised by the caller (one host thread per GPU, as in the reference).
 */
#ifndef TFB_H
#define TFB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFB_ABI_VERSION 1

typedef enum {
    TFB_OK = 0,
    TFB_ERR_CUDA = -1,        /* a CUDA runtime call failed            */
    TFB_ERR_ARG = -2,         /* bad argument (NULL, size, flag combo)  */
    TFB_ERR_NO_DEVICE = -3,   /* no CUDA device / wrong architecture    */
    TFB_ERR_NOT_CONVERGED = -4
} tfb_status;

/* scalar-or-grid fp64 input */
typedef struct {
    const double *ptr;        /* device pointer to (ny, nx) grid, or NULL */
    double val;               /* used when ptr == NULL                    */
} tfb_sg;

int tfb_abi_version(void);
const char *tfb_last_error(void);
/* Block the calling host thread until `stream` has drained (cudaStreamSynchronize): how the
 * component waits for the scalar block a step wrote into pinned host memory -- the
 * reference reads Q_outlet etc. right after update() (channels_base.py:2884-2936). */
int tfb_stream_synchronize(void *stream);
/* name, SM count, cc major/minor of the current device */
int tfb_device_info(char *name, int name_len, int *sm_count, int *cc_major,
                    int *cc_minor);

/* ------------------------------------------------------------------------
 * D8 toolkit  (components/d8_base.py, components/d8_global.py)
 * ---------------------------------------------------------------------- */

/* 256-entry tie-break table, copied to `out256` (HOST pointer).
 * replaces d8_base.get_resolve_array (d8_base.py:711-772). */
int tfb_d8_resolve_table(uint8_t *out256);

/* D8 start codes from a float32 DEM: steepest-descent slopes to the 8
 * neighbours in fp32, ties encoded as the SUM of tied codes, flats negated
 * (link_flats != 0) or zeroed, pits -300 / 0, nodata/NaN and the four border
 * lines 0.  `dem` points at the first OWNED row of a slab; `halo` rows above
 * and below it are readable (0 for a whole grid: its border rows get code 0
 * anyway); `row0`/`ny_global` place the slab in the global grid.
 * replaces d8_global.start_new_d8_codes (d8_global.py:181-353). */
int tfb_d8_codes(const float *dem, int32_t ny, int32_t nx, int32_t halo,
                 int32_t row0, int32_t ny_global, float dx, float dy, float dd,
                 double nodata, int32_t link_flats, int16_t *code, void *stream);

/* code[code > 0] = resolve[code]   (d8_global.break_flow_grid_ties :355-372) */
int tfb_d8_break_ties(int16_t *code, int64_t n, void *stream);

/* ONE Jacobi sweep of flat linking: reads `code_in` (same halo convention),
 * writes `code_out`; n_flats / n_fixed are DEVICE int32 counters that the sweep
 * ADDS to.  (d8_global.link_flats :374-576, one trip of its while loop) */
int tfb_d8_link_flats_sweep(const int16_t *code_in, int16_t *code_out,
                            int32_t ny, int32_t nx, int32_t halo, int32_t row0,
                            int32_t ny_global, int32_t *n_flats,
                            int32_t *n_fixed, void *stream);

/* Whole link_flats loop on one GPU (iterates sweeps to the reference's stop
 * rule).  `scratch` = device int16 buffer of ny*nx.  n_reps: HOST out. */
int tfb_d8_link_flats(int16_t *code, int16_t *scratch, int32_t ny, int32_t nx,
                      int32_t max_reps, int32_t *n_reps, void *stream);

/* negative / >128 leftovers -> 0, then to uint8 (d8_global.py:139-146) */
int tfb_d8_finalize_codes(const int16_t *code, uint8_t *code_u8, int64_t n,
                          void *stream);

/* ds (flow length) and dw (flow width) float32 grids from the codes and the
 * per-row float32 pixel sizes dx, dy, dd (DEVICE arrays of ny).
 * (d8_global.update_flow_width_grid :873-958, update_flow_length_grid :960-1027)
 * Either output may be NULL. */
int tfb_d8_ds_dw(const uint8_t *code, const float *dx, const float *dy,
                 const float *dd, int32_t ny, int32_t nx, float *ds, float *dw,
                 void *stream);

/* Total contributing area: A[cell] = pixel_area (stored fp32) + sum of the
 * FINAL areas of the children, added in neighbour order NE,E,SE,S,SW,W,NW,N in
 * fp32 -- the reference's result, computed in one topological pass instead of
 * one whole-grid sweep per level.  Cells with code 0 and cells on / below a
 * cycle keep A = 0.  `pixel_area_row` (DEVICE, ny doubles) may be NULL, then
 * `pixel_area` applies.  `count` (DEVICE int64, may be NULL) receives the exact
 * number of cells draining through each cell (itself included).
 * `pending` = DEVICE scratch of ny*nx int32.   n_done: HOST out, number of
 * cells that received an area.   (d8_global.update_area_grid :1029-1332) */
int tfb_d8_area(const uint8_t *code, int32_t ny, int32_t nx, double pixel_area,
                const double *pixel_area_row, float *A, int64_t *count,
                int32_t *pending, int64_t *n_done, void *stream);

/* Row-slab form of tfb_d8_area (one process per GPU; d8_global.update_area_grid
 * :1029-1332 on a grid sharded by rows).  `code`, `cell` (uint64 words {fp32 area,
 * uint32 count}) and `word` (int32 scratch) point at the first OWNED row of arrays
 * that carry ONE halo row above and below.  _init clears the words and analyses the
 * slab; _walk with resume = 0 runs the topological walk from the slab's own source
 * cells -- a walker stops where its next cell belongs to a neighbour; the caller then
 * exchanges the boundary rows of `cell` into the neighbours' halo rows (2 x nx x 8 B)
 * and calls _walk with resume = 1, which lets every halo cell whose word has arrived
 * continue into this slab; repeat until no boundary row changes on any rank.
 * tfb_d8_area_unpack splits the words of n cells into A (float32) and count (int64,
 * may be NULL); n_done (HOST, may be NULL) = cells that received an area.  Areas and
 * counts are bit-identical to the single-GPU pass. */
int tfb_d8_area_slab_init(const uint8_t *code, int32_t ny, int32_t nx, int32_t row0,
                          int32_t ny_global, uint64_t *cell, int32_t *word,
                          void *stream);
int tfb_d8_area_slab_walk(int32_t ny, int32_t nx, double pixel_area,
                          const double *pixel_area_row, uint64_t *cell,
                          const int32_t *word, int32_t resume, void *stream);
int tfb_d8_area_unpack(const uint64_t *cell, int64_t n, float *A, int64_t *count,
                       int64_t *n_done, void *stream);

/* S = (DEM - DEM[parent]) / ds in fp32, 0 at noflow cells
 * (d8_global.update_slope_grid :1334-1348). */
int tfb_d8_slope(const float *dem, const uint8_t *code, const float *dx,
                 const float *dy, const float *dd, int32_t ny, int32_t nx,
                 float *S, void *stream);

/* D8 "aspect": flow angle in radians, counter-clockwise from east, float32, 0
 * for noflow cells (d8_global.update_aspect_grid :1350-1376). */
int tfb_d8_aspect(const uint8_t *code, int64_t n, float *aspect, void *stream);

/* parent_ID_grid (int32 flat IDs, 0 for noflow/border cells)
 * (d8_global.update_parent_ID_grid :578-650). */
int tfb_d8_parent_ids(const uint8_t *code, int32_t ny, int32_t nx,
                      int32_t *pid, void *stream);

/* ------------------------------------------------------------------------
 * Channel routing step  (components/channels_base.py update() :803-962 and
 * subclasses channels_kinematic_wave.py :84-127, channels_diffusive_wave.py
 * :85-139) with optionally fused source terms.
 * ---------------------------------------------------------------------- */

/* flags */
#define TFB_CH_DIFFUSIVE   (1u << 0)  /* S = S_free (else kinematic: S_bed)   */
#define TFB_CH_LAW_OF_WALL (1u << 1)  /* velocity law (else Manning)          */
#define TFB_CH_DIAG        (1u << 2)  /* materialise Rh,A_wet,P_wet,f,tau,
                                         u_star,froude (+S_free) every step   */
#define TFB_CH_WRITE_R     (1u << 3)  /* store the R grid                     */
#define TFB_CH_ET_INPLACE  (1u << 4)  /* zero the provider's ET grid where
                                         vol < ET*da*dt (channels_base :1625)  */
#define TFB_CH_FLOOD       (1u << 5)  /* FLOOD_OPTION: rectangle-over-trapezoid
                                         overbank flow (channels_base.py:1701-1862,
                                         :2149-2185, :2393-2448); generic kernel only */
#define TFB_CH_ATTENUATE   (1u << 6)  /* ATTENUATE: runoff passes through a per-cell
                                         linear reservoir (channels_base.py:2077-2084);
                                         generic kernel only                        */
#define TFB_CH_SRC_MET     (1u << 8)  /* P -> P_rain/P_snow split fused       */
#define TFB_CH_SRC_SNOW    (1u << 9)  /* degree-day snowmelt fused            */
#define TFB_CH_SRC_EVAP    (1u << 10) /* Priestley-Taylor ET fused            */
#define TFB_CH_SRC_INFIL   (1u << 11) /* Green-Ampt infiltration fused        */
#define TFB_CH_REVERSE     (1u << 12) /* packed kinematic step: walk the tiles bottom-up.  A
                                         caller that alternates the bit makes every step start
                                         on the rows the previous one wrote LAST, while they
                                         still sit in the 126 MB L2.  Per-cell results do not
                                         depend on it; the grid sums associate differently
                                         (deterministically for a given bit).  Diffusive wave:
                                         set it for pass B only (A top-down, B bottom-up: B
                                         starts on the depths A wrote last, the next A on the
                                         discharges B wrote last); pass A ignores it           */

/* Device-resident scalar block: what the reference keeps in 0-D numpy arrays.
 * The step kernel's last block updates it; the host reads it when it needs
 * Q_outlet etc. (one D2H of sizeof(tfb_channels_scalars)). */
typedef struct {
    double vol_R, vol_Q, vol_edge;                 /* :1657, :2928, :2962     */
    double Q_outlet, u_outlet, d_outlet, f_outlet; /* :2884-2906              */
    double Q_peak, T_peak, u_peak, Tu_peak, d_peak, Td_peak; /* :2908-2926    */
    double vol_P, vol_PR, vol_PS;                  /* met_base.py:1247-1446   */
    double vol_SM, vol_ET, vol_IN;                 /* snow/evap/infil integrals */
    double dt_cfl_min;                             /* diagnostic only: min ds/u */
    int64_t n_neg_d, n_nan_d, n_inf_d;             /* check_flow_depth :3193  */
    int64_t n_neg_u, n_nan_u, n_inf_u;             /* check_flow_velocity :3352 */
    int64_t n_nan_IN;                              /* infil_base.py:906-950   */
    int64_t step_count;
    int64_t n_step_failed;    /* launches (cumulative) in which a bounded wait of the TMA
                                 pipeline or of the peer-memory halo timed out: the grids
                                 of such a step are NOT valid; the host raises on it    */
} tfb_channels_scalars;

typedef struct {
    /* ---- geometry ---------------------------------------------------- */
    int32_t ny, nx;           /* rows / cols held by this GPU (its slab)        */
    int32_t row0, ny_global;  /* slab position in the global grid               */
    uint32_t flags;
    int32_t outlet_row, outlet_col;  /* GLOBAL indices; sampled by the owner    */
    int32_t is_R_grid;        /* 1: vol_R sums R*da*dt per cell; 0: scalar rule  */
    double dt;                /* channel time step [s]                          */
    double time_min;          /* model time [min] at the start of the step      */
    double rho_H2O;
    int64_t n_pixels_global;  /* for scalar-R integral (channels_base :1668)    */
    /* ---- static per-cell inputs --------------------------------------- */
    int32_t halo;             /* rows readable above row 0 and below row ny-1 of
                                 EVERY grid pointer (0, or >=1 KW / >=2 DW for a
                                 slab with neighbours); rows beyond that read as
                                 code 0 / Q 0                                   */
    int32_t pitch;            /* elements between the starts of consecutive rows of
                                 EVERY per-cell (ny, nx) grid pointer of this struct
                                 and of tfb_channels_packed (0 = nx: dense rows).
                                 The TMA path needs 16-byte row strides, so a caller
                                 with nx % 4 != 0 (Treynor: nx = 29) allocates rows
                                 of pitch = nx rounded up to 4 elements; columns
                                 [nx, pitch) are never read as cells nor written.
                                 Per-ROW arrays (dx, dy, dd, da.ptr, row_geom) are
                                 not affected                                      */
    const uint16_t *geo;      /* packed per-cell word from tfb_channels_pack_geo:
                                 low byte D8 code, high byte child mask          */
    const float *dx, *dy, *dd;        /* per-row pixel sizes (ny), float32       */
    tfb_sg sinu;              /* ds = sinu * ds_d8  (channels_base :1242)       */
    tfb_sg da;                /* pixel area [m2]; ptr = per-ROW array (ny)      */
    const double *S_bed;      /* bed slope (already / sinu, bad slopes fixed);
                                 needed for DIFFUSIVE, LAW_OF_WALL or DIAG      */
    const double *sqrtS;      /* sqrt(|S_bed|) evaluated on the host: the only
                                 use of S_bed in the lean kinematic Manning step */
    tfb_sg width;             /* trapezoid bottom width                         */
    tfb_sg tan_angle;         /* tan / cos of the bank angle, evaluated on the  */
    tfb_sg cos_angle;         /*   host by numpy at initialize(); the reference's
                                 angle == 0 test is tan_angle == 0              */
    tfb_sg rough;             /* Manning n  or  z0 (law of the wall)            */
    /* ---- state --------------------------------------------------------- */
    const double *vol;        /* in: water volume per cell before the step      */
    double *vol_out;          /* out: after the step; may alias `vol` for the
                                 kinematic wave, must NOT for the diffusive one
                                 (neighbours re-read the old volumes)           */
    const double *Q_in;       /* discharge used by this step (= u*A_wet of the
                                 previous one)                                  */
    double *Q_out;            /* out: u*A_wet after this step                   */
    double *d, *u;            /* out                                            */
    /* ---- optional outputs (TFB_CH_DIAG / TFB_CH_WRITE_R), may be NULL --- */
    double *Rh, *A_wet, *P_wet, *f, *tau, *u_star, *froude, *S_free, *R;
    /* ---- source terms --------------------------------------------------- */
    tfb_sg P_rain;            /* or total P when TFB_CH_SRC_MET                 */
    tfb_sg SM, GW, MR, IN;    /* used as given unless the fused flag is set     */
    tfb_sg ET;
    double *ET_rw;            /* == ET.ptr when TFB_CH_ET_INPLACE               */
    /* met (met_base.py:1300-1388) */
    tfb_sg T_air; double T_rain_snow;
    /* snow (snow_degree_day.py:293-360, snow_base.py:490-703) */
    tfb_sg c0, T0; const double *h_swe; double *h_swe_out; /* alias rule as vol */
    double *h_snow; double rho_snow;
    /* evap (evap_priestley_taylor.py:308-413, evap_base.py:326-353) */
    tfb_sg T_surf, Qn_SW, Qn_LW; double alpha, K_soil, soil_x, T_soil_x;
    /* infil (infil_green_ampt.py:511-578, infil_base.py:654-811, :954-995) */
    tfb_sg Ks, Ki, qs, qi, G; const double *I; double *I_out; /* alias rule as vol */
    tfb_sg h_table, elev;
    /* ---- reductions ----------------------------------------------------- */
    tfb_channels_scalars *scalars;   /* DEVICE                                  */
    double *partials;         /* DEVICE scratch, >= tfb_channels_partials_len()
                                 doubles, zero-initialised once by the caller;
                                 after a step its FIRST TFB_NSUM doubles hold
                                 this rank's sums                               */
    uint32_t *ticket;         /* DEVICE scratch, 2 words, zero-initialised      */
    int32_t finalize;         /* 1: last block folds partials into `scalars`
                                 (single GPU); 0: only leave the per-rank sums
                                 in partials[0..TFB_NSUM) for an allreduce, then
                                 call tfb_channels_fold()                       */
    tfb_channels_scalars *scalars_mirror;  /* optional: device-visible address of
                                 PINNED HOST memory; the folding thread copies the
                                 updated scalar block there (posted writes over
                                 PCIe), so the host reads Q_outlet etc. after a
                                 stream synchronise without issuing a copy        */
    /* ---- FLOOD_OPTION (TFB_CH_FLOOD) ------------------------------------------ */
    tfb_sg vol_bankfull;      /* channel volume at bankfull depth:
                                 d_b (w + d_b tan(angle)) ds   (:1385-1387)      */
    double *d_flood;          /* out: flood-plain water depth (may be NULL)      */
    double *vol_flood;        /* out: water volume above bankfull (may be NULL)  */
    double flood_manning_n;   /* 0.10 in the reference (:1367)                   */
    double width_ratio;       /* flood-plain / channel width, 5.0 (:1368)        */
    /* ---- ATTENUATE (TFB_CH_ATTENUATE) ------------------------------------------ */
    const double *vol_stored; /* in: water held back on the hillslope             */
    double *vol_stored_out;   /* out (alias rule as vol)                          */
    double t_drain;           /* 86400 * n_days [s]                               */
} tfb_channels_step_args;

#define TFB_NSUM 24           /* doubles per rank exchanged by the allreduce:
                                 [0,15) sums and bad-value counts, [15] "this
                                 rank owns the outlet", [16] dt_cfl_min (reduce
                                 with MIN or ignore), [17,21) outlet Q,u,d,f
                                 (zero on non-owners), [22] CTAs whose bounded
                                 waits timed out, rest padding                  */

/* sizeof() of the two structs above, so a binding can verify its own layout */
int64_t tfb_sizeof_channels_step_args(void);
int64_t tfb_sizeof_channels_scalars(void);

/* number of doubles `partials` must hold for a (ny, nx) slab */
int64_t tfb_channels_partials_len(int32_t ny, int32_t nx);

/* geo word of every cell of a slab INCLUDING its halo rows: code | mask << 8,
 * mask bit k set <=> the neighbour that would drain into the cell with code
 * 1<<k has that code (the per-direction w_k / p_k lists of
 * d8_global.update_flow_from_IDs :723-764 / update_flow_to_IDs :766-841 as a
 * parent-side gather mask).  `code` / `geo` point at the first owned row. */
int tfb_channels_pack_geo(const uint8_t *code, int32_t ny, int32_t nx,
                          int32_t halo, uint16_t *geo, void *stream);

/* one routing step of the slab */
int tfb_channels_step(const tfb_channels_step_args *a, void *stream);

/* ---- packed fast path ------------------------------------------------------
 * The common production case -- kinematic wave, Manning, no per-step
 * diagnostics, scalar sinuosity and pixel area, scalar forcing, optionally
 * fused Green-Ampt with scalar soil parameters -- runs from a COMPACT, lossless
 * re-encoding of the static per-cell inputs built once at initialize():
 *   geo32     code | child mask << 8 | bank-angle index << 16   (4 B / cell)
 *   width32   bottom width as float32 (the RTG files hold float32; the caller
 *             must have verified (double)(float)w == w)          (4 B / cell)
 *   coef      sqrt(|S_bed|) / n  (both static, channels_base.py:3982-3983)
 *                                                                 (8 B / cell)
 *   angle_lut per distinct bank angle (<= 256): tan, 1/(2 tan) (0 when tan==0),
 *             1/cos, 4 tan
 * i.e. 16 B of static input per cell instead of 42.  Divisions by static
 * quantities (ds, 2 tan, cos, n) become multiplications by reciprocals that are
 * themselves correctly rounded, so results differ from the exact-order generic
 * kernel by a few ulp per step (parity tests: rel <= 1e-11 over a storm).
 * All pointers address the first OWNED row of (ny, nx) arrays. */
typedef struct {
    const uint32_t *geo32;
    const float *width32;
    const double *coef;
    const double *angle_lut;  /* n_angles x 4 doubles, DEVICE */
    const double *row_geom;   /* ny x 8 doubles from tfb_channels_pack_rows: flow
                                 length ds = sinu * {dx, dd, dy} of the row (E/W and
                                 noflow, diagonal, N/S), the 3 reciprocals, the pixel
                                 area of the row, padding                        */
    double outlet_rough;      /* Manning n of the outlet cell (f_outlet only)    */
    int32_t n_angles;
    /* diffusive wave only (coef may then be NULL): */
    const double *inv_rough;  /* 1 / n                                   (8 B / cell) */
    const float *S32;         /* bed slope as float32 (lossless)         (4 B / cell) */
    const int64_t *noflow_idx; /* flat indices (owned rows, row pitch) of the cells with code 0 */
    int64_t n_noflow;
    /* extended variants (tfb_channels_step_ext*): */
    const double *rough64;    /* Manning n, or z0 for the law of the wall  (8 B / cell) */
    const double *lut_tc;     /* n_angles x 2 doubles: tan, cos of every distinct bank
                                 angle; geo32 carries a 16-bit index, so up to 65536
                                 angles (the lean kernels keep their <= 256-entry
                                 angle_lut in shared memory)                           */
    /* extended variants, inputs that are NOT float32 values (a scalar width such as 3.3 m
     * from the CFG; slopes divided by a sinuosity != 1, channels_base.py:1247): the fp64
     * grid takes the place of width32 / S32 (which may then be NULL)        (8 B / cell) */
    const double *width64;
    const double *S64;
    const double *sinu64;     /* sinuosity GRID (ds = sinu * ds_d8 per cell, :1242);
                                 row_geom is then built with sinu = 1         (8 B / cell) */
} tfb_channels_packed;

int64_t tfb_sizeof_channels_packed(void);

/* geo32 of the owned rows from the D8 codes (halo rows readable as in
 * tfb_channels_pack_geo) and a per-cell bank-angle index (owned rows) */
int tfb_channels_pack_geo32(const uint8_t *code, const uint16_t *angle_idx,
                            int32_t ny, int32_t nx, int32_t halo,
                            uint32_t *geo32, void *stream);

/* per-row flow lengths and their reciprocals (d8_global.py:960-1027 times the
 * scalar sinuosity, channels_base.py:1242) and the pixel area of the row
 * (utils/pixels.py:173-233); dx, dy, dd = DEVICE float32 (ny); da_row = DEVICE
 * float64 (ny) or NULL, then `da` applies to every row */
int tfb_channels_pack_rows(const float *dx, const float *dy, const float *dd,
                           double sinu, const double *da_row, double da,
                           int32_t ny, double *row_geom, void *stream);

/* one routing step through the packed path; TFB_ERR_ARG when the arguments do
 * not meet the conditions above (the caller then uses tfb_channels_step).
 * Reads a->vol, a->Q_in (with halo), a->I; writes a->vol_out, d, u, Q_out,
 * I_out; uses a->partials / ticket / scalars like tfb_channels_step. */
int tfb_channels_step_packed(const tfb_channels_step_args *a,
                             const tfb_channels_packed *p, void *stream);

/* Diffusive wave through the packed path: two launches per step (+ a tiny third).  Pass A updates
 * the volumes and writes the NEW depth of every cell (update_flow_volume :1951,
 * update_channel_depth :2250, fused sources, volume integrals); a slab run then
 * exchanges one boundary row of `d`; pass B forms the free-surface slope from the
 * new depths of each cell and its D8 parent (update_free_surface_slope
 * :2583-2618), the velocity and the next discharge.  Same eligibility rules as
 * tfb_channels_step_packed. */
int tfb_channels_step_packed_dw_a(const tfb_channels_step_args *a,
                                  const tfb_channels_packed *p, void *stream);
int tfb_channels_step_packed_dw_b(const tfb_channels_step_args *a,
                                  const tfb_channels_packed *p, void *stream);
/* third, tiny launch of the diffusive step: d = 0 at the noflow cells
 * (update_edge_values :2966).  It cannot be part of pass B, whose tiles read the
 * PRE-zero depth of noflow parents in neighbouring tiles (:2583-2618 runs before
 * :2938 in the reference). */
int tfb_channels_step_packed_dw_c(const tfb_channels_step_args *a,
                                  const tfb_channels_packed *p, void *stream);

/* ---- extended packed step -----------------------------------------------------
 * The same TMA-ring kernels for what a COUPLED run delivers and the lean variants
 * above reject: any of P_rain, SM, GW, MR, ET, IN as (ny, nx) grids that change
 * every step (update_R, channels_base.py:1548-1655; the provider's ET grid is
 * masked in place through a->ET_rw, :1625-1627; a->R is written with
 * TFB_CH_WRITE_R), TFB_CH_LAW_OF_WALL (:4005-4063), TFB_CH_DIAG (all seven
 * diagnostic grids, + S_free in pass B), up to 65536 distinct bank angles.  Needs
 * p->S32, p->rough64, p->lut_tc besides geo32 / width32 / row_geom; no fused
 * sources, flood option or ATTENUATE (those stay with tfb_channels_step).  The
 * arithmetic keeps the reference's operation order (IEEE divisions), unlike the
 * lean variants' reciprocal forms.  Diffusive wave: _dw_a, (slab: exchange one row
 * of d), _dw_b, then tfb_channels_step_packed_dw_c as in the lean path. */
int tfb_channels_step_ext(const tfb_channels_step_args *a,
                          const tfb_channels_packed *p, void *stream);
int tfb_channels_step_ext_dw_a(const tfb_channels_step_args *a,
                               const tfb_channels_packed *p, void *stream);
int tfb_channels_step_ext_dw_b(const tfb_channels_step_args *a,
                               const tfb_channels_packed *p, void *stream);

/* ---- peer-memory halo (NVLink / NVSwitch) ----------------------------------
 * One process per GPU; the slabs' discharge buffers are allocated with
 * tfb_peer_alloc, their CUDA-IPC handles exchanged by the host, and the
 * neighbours' halo rows mapped with tfb_peer_open.  tfb_channels_step_packed_p2p
 * then fuses the halo exchange into the routing kernel: the warps that compute
 * the first / last owned row store the new discharge ALSO into the neighbour's
 * halo row (peer stores over NVLink).  The first and last tile rows are
 * processed FIRST; when the last of them has retired one thread publishes
 * "step k done" in a flag word in the neighbour's memory, so the transfer
 * overlaps the rest of the grid.  On the receiving side the producer warp of a
 * boundary tile polls its local flag (ld.acquire.sys, bounded) until the
 * neighbour has published the previous step -- which orders both the halo reads
 * and the re-use of the double-buffered halo rows -- before it issues that
 * tile's TMA loads; in lock-step operation the flag has been set for almost a
 * whole step by then.  No NCCL call, no stream memory operation and no extra
 * kernel per step (TFB_P2P_STREAM_WAIT=1 in the environment selects
 * cuStreamWaitValue32 in front of the launch instead). */
typedef struct {
    double *up_Q;             /* neighbour above: its LOWER halo row of the buffer this
                                 step writes (row ny_up of its owned rows), or NULL   */
    double *down_Q;           /* neighbour below: its UPPER halo row (row -1), or NULL */
    uint32_t *up_flag;        /* word in the neighbour's memory: steps I have completed */
    uint32_t *down_flag;
    const uint32_t *from_up;  /* LOCAL words the neighbours write (NULL: no neighbour) */
    const uint32_t *from_down;
    uint32_t step;            /* 0-based index of this step                            */
} tfb_peer_halo;

int64_t tfb_sizeof_peer_halo(void);
/* cudaMalloc + zero fill + IPC handle (64 bytes) of a buffer peers may map */
int tfb_peer_alloc(int64_t bytes, void **ptr, uint8_t *handle64);
int tfb_peer_open(const uint8_t *handle64, void **ptr);
int tfb_peer_close(void *ptr);
int tfb_peer_free(void *ptr);
int tfb_channels_step_packed_p2p(const tfb_channels_step_args *a,
                                 const tfb_channels_packed *p,
                                 const tfb_peer_halo *peer, void *stream);
/* The two passes of the diffusive step with the same peer-memory halo.  `step` is a
 * PHASE counter here: 2n for pass A of step n (waits for the neighbours' 2n =
 * their pass B of step n-1 has read its depth halo and stored my discharge halo;
 * publishes 2n+1), 2n+1 for pass B (waits for 2n+1 = the neighbours' pass A
 * stored my depth halo; publishes 2n+2).  up_Q / down_Q address the neighbours'
 * halo rows of the array the pass WRITES: the depth `d` (pre-zero values) for
 * pass A, the discharge buffer for pass B.  d and both discharge buffers must
 * therefore come from tfb_peer_alloc. */
int tfb_channels_step_packed_dw_a_p2p(const tfb_channels_step_args *a,
                                      const tfb_channels_packed *p,
                                      const tfb_peer_halo *peer, void *stream);
int tfb_channels_step_packed_dw_b_p2p(const tfb_channels_step_args *a,
                                      const tfb_channels_packed *p,
                                      const tfb_peer_halo *peer, void *stream);

/* multi-GPU: fold globally reduced sums (DEVICE, TFB_NSUM doubles laid out as
 * tfb_channels_step leaves them in partials[]) into the scalar block */
int tfb_channels_fold(const tfb_channels_step_args *a, const double *sums,
                      void *stream);

/* domain min/max of Q, u, d over interior cells and total channel volume
 * (channels_base.update_mins_and_maxes :2986-3081,
 *  update_total_channel_water_volume :3083-3110); out6 = DEVICE
 * {Q_min,Q_max,u_min,u_max,d_min,d_max}, vol_sum = DEVICE */
/* doubles of scratch the stand-alone source kernels and tfb_channels_minmax
 * need in `partials` (their `ticket` is one zero-initialised uint32) */
int64_t tfb_sources_partials_len(void);

int tfb_channels_minmax(const double *Q, const double *u, const double *d,
                        const double *vol, int32_t ny, int32_t nx, double *out6,
                        double *vol_sum, double *partials, uint32_t *ticket,
                        void *stream);

/* ------------------------------------------------------------------------
 * Stand-alone source-term kernels (one per reference component update()).
 * ---------------------------------------------------------------------- */

/* P_rain = P*(T_air > T_rs), P_snow = P*(T_air <= T_rs); sums[0..3) +=
 * {vol_P, vol_PR, vol_PS} increments.  (met_base.py:1247-1446) */
int tfb_met_precip(tfb_sg P, tfb_sg T_air, double T_rain_snow, tfb_sg da,
                   double dt, int32_t ny, int32_t nx, double *P_rain,
                   double *P_snow, double *sums3, double *partials,
                   uint32_t *ticket, void *stream);

/* Green-Ampt v1 (infil_green_ampt.py:511-578; infil_base.update :301-357):
 * IN out, I in/out, sums[0] += vol_IN increment, nbad += #non-finite IN */
int tfb_infil_green_ampt(tfb_sg P_rain, tfb_sg SM, tfb_sg Ks, tfb_sg Ki,
                         tfb_sg qs, tfb_sg qi, tfb_sg G, tfb_sg h_table,
                         tfb_sg elev, int32_t use_table, tfb_sg da, double dt,
                         int32_t ny, int32_t nx, double *IN, double *I,
                         double *sums2, double *partials, uint32_t *ticket,
                         void *stream);

/* degree-day (snow_degree_day.py:293-360; snow_base.update :214-278) */
int tfb_snow_degree_day(tfb_sg P_snow, tfb_sg T_air, tfb_sg c0, tfb_sg T0,
                        double rho_H2O, double rho_snow, tfb_sg da, double dt,
                        int32_t ny, int32_t nx, double *SM, double *h_swe,
                        double *h_snow, double *sums1, double *partials,
                        uint32_t *ticket, void *stream);

/* Priestley-Taylor (evap_priestley_taylor.py:308-413; evap_base :326-377) */
int tfb_evap_priestley_taylor(tfb_sg T_air, tfb_sg T_surf, tfb_sg Qn_SW,
                              tfb_sg Qn_LW, double alpha, double K_soil,
                              double soil_x, double T_soil_x, tfb_sg da,
                              double dt, int32_t ny, int32_t nx, double *Qc,
                              double *ET, double *sums1, double *partials,
                              uint32_t *ticket, void *stream);

/* Meteorology energy chain (met_base.update() :790-809 with PRECIP_ONLY = No):
 * vapour pressures, dew point, surface temperature, bulk Richardson number and
 * aerodynamic conductances, sensible / latent heat, precipitable water,
 * clear-sky shortwave on a sloping cell (solar_funcs.Clear_Sky_Radiation :870 and
 * everything below it), air emissivity, net longwave, net energy flux.  Every
 * input is scalar-or-grid; outputs are optional DEVICE grids.  julian_day (and
 * TSN_offset, the clock time relative to true solar noon [h]) are formed on the
 * host from the model date (met_base.update_julian_day :1938-1993).  Albedo is
 * an input here; tfb_met_albedo updates it. */
typedef struct {
    int32_t ny, nx;
    int32_t satterlund;       /* saturation vapour pressure method (else Brutsaert) */
    tfb_sg T_air, RH, uz, z, z0_air, h_snow, h_ice, p0, albedo, em_surf;
    tfb_sg lat_deg, alpha, beta, TSN_offset, cloud_factor, canopy_factor, Qa, Qc;
    double julian_day, dust_atten;
    double *e_sat_air, *e_air, *T_dew, *T_surf, *e_sat_surf, *Ri, *Dn, *Dh, *Qh, *W_p,
           *e_surf, *Qe, *Qn_SW, *em_air, *Qn_LW, *Q_sum;
} tfb_met_energy_args;

/* Albedo (met_base.update_albedo :1995-2045), in place on the DEVICE grid `albedo`.
 * aging != 0: snow albedo 0.4 + 0.44 exp(-n r) with n = days since the snowfall of
 * the last three days reached 3 cm; `ring` = DEVICE doubles [slots][ny*nx] holding
 * the per-step snow depths P_snow*dt*rho_H2O/rho_snow (slots = int(259200/dt),
 * met_base.py:1126-1133), `head` = slot that receives this step's value (the
 * caller advances it by one, modulo slots, per call), `n_days` = DEVICE grid.
 * aging == 0: the 'simple' method (0.75 on snow).  Both: 0.3 on snow-free ice,
 * 0.15 on bare ground. */
int tfb_met_albedo(tfb_sg P_snow, tfb_sg T_air, tfb_sg h_snow, tfb_sg h_ice,
                   int32_t ny, int32_t nx, double dt, double rho_H2O, double rho_snow,
                   int32_t aging, double *ring, int32_t slots, int32_t head,
                   double *n_days, double *albedo, void *stream);

int64_t tfb_sizeof_met_energy_args(void);
int tfb_met_energy(const tfb_met_energy_args *a, void *stream);

/* ------------------------------------------------------------------------
 * DEM / channel-geometry preparation.
 * ---------------------------------------------------------------------- */

/* Depression filling, IN PLACE on the float32 DEM: utils/fill_pits.py
 * fill_pits() :358-640 (priority flood from the seeds of get_start_pixels()
 * :178-356: valid border cells, cells next to nodata, cells holding
 * closed_code; non-finite values become `nodata`).  The result is the unique
 * lowest surface >= DEM that drains to the seeds, computed by tiled chaotic
 * relaxation (bit-identical to the heap order of the reference).
 * close_nodata = 0: what the reference does -- :404-409 uses the ROW numbers of
 * the nodata cells as flat cell IDs, so nodata cells stay open and are flooded
 * up to their rim while cells with flat index r < ny (r a row holding nodata)
 * are walled off; close_nodata = 1: the documented intent (nodata cells closed).
 * `scratch` = DEVICE, tfb_fill_pits_scratch_bytes(ny, nx) bytes.  HOST outputs:
 * n_raised (cells whose elevation was raised), n_rounds (relaxation rounds). */
int64_t tfb_fill_pits_scratch_bytes(int32_t ny, int32_t nx);
int tfb_fill_pits(float *dem, int32_t ny, int32_t nx, float nodata, float closed_code,
                  int32_t close_nodata, void *scratch, int64_t *n_raised, int32_t *n_rounds, void *stream);

/* slope = (z - z[first downstream cell that is lower]) / path length, fp64:
 * utils/new_slopes.py get_new_slope_grid() :174-316, including its treatment of
 * walkers that run past their noflow cell (parent ID 0 = cell (0, 0)).
 * `scratch` = DEVICE int32[2*ny*nx + 1]; longest_path (HOST, may be NULL) = the
 * number of iterations the reference's loop executes. */
int tfb_d8_new_slope(const float *dem, const uint8_t *code, const float *dx,
                     const float *dy, const float *dd, int32_t ny, int32_t nx,
                     float nodata, int32_t *scratch, double *slope,
                     int32_t *longest_path, void *stream);

/* smallest positive value and maximum of a float32 grid (HOST outputs; +inf / 0
 * when no value is positive) -- the A1, A2 of get_grid_from_TCA(). */
int tfb_area_range(const float *A, int64_t n, float *min_pos, float *max_val, void *stream);

/* grid = c * A^p, 1.0 where A <= 0: utils/parameterize.py get_grid_from_TCA()
 * :93-215 (channel width, Manning n, bankfull depth from contributing area).
 * single_precision != 0 rounds A^p and the product to float32 as numpy does
 * when p and c are float32 there (the "p and g1 given" branch).  Either output
 * may be NULL. */
int tfb_power_law_grid(const float *A, int64_t n, double c, double p,
                       int32_t single_precision, double *out64, float *out32, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TFB_H */
