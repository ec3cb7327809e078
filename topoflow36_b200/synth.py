"""Seeded synthetic cases: a tilted fractal DEM, channel
parameter grids, and a writer that lays a case out on disk in the reference's
own formats (RTI + RTG + pipe-delimited CFG) so the SAME files drive the
reference, the oracle and the GPU components.
"""
import os

import numpy as np

from . import grid_io


def fractal_dem(ny, nx, seed=1234, dx=30.0, dy=30.0, sy=0.02, sx=0.005,
                beta=2.2, sigma=2.0, z0=None):
    """z(r,c) = z0 - sy*dy*r - sx*dx*c + F(r,c), float32; F = spectral-synthesis
    noise with power spectrum ~ k^-beta scaled to `sigma` metres."""
    rng = np.random.default_rng(seed)
    white = rng.standard_normal((ny, nx))
    ky = np.fft.fftfreq(ny)[:, None]
    kx = np.fft.rfftfreq(nx)[None, :]
    k = np.sqrt(ky * ky + kx * kx)
    k[0, 0] = 1.0
    amp = k ** (-beta / 2.0)
    amp[0, 0] = 0.0
    F = np.fft.irfft2(np.fft.rfft2(white) * amp, s=(ny, nx))
    F *= sigma / F.std()
    r = np.arange(ny, dtype=np.float64)[:, None]
    c = np.arange(nx, dtype=np.float64)[None, :]
    if z0 is None:
        z0 = sy * dy * ny + sx * dx * nx + 10.0 * sigma + 100.0
    return np.float32(z0 - sy * dy * r - sx * dx * c + F)


def channel_params(ny, nx, seed=1234):
    """width [m], bank angle [deg], Manning n -- float32 as they sit on disk."""
    rng = np.random.default_rng(seed + 1)
    width = np.float32(1.0 + 4.0 * rng.random((ny, nx)))
    angle = np.full((ny, nx), 45.0, dtype=np.float32)
    nval = np.float32(0.03 + 0.02 * rng.random((ny, nx)))
    return width, angle, nval


_CHAN_CFG = """#===============================================================================
# TopoFlow Config File for: Channels_{title}
#===============================================================================
# Input 1
comp_status         | Enabled      | string    | component status {{Enabled; Disabled}}
stop_method         | Until_n_steps | string    | stopping method
Qp_fraction         | 0.05      | float     | Value for Q_peak_fraction method
T_stop_model        | 2000     | float     | Value for Until_model_time method [minutes]
n_steps             | {n_steps}         | int       | Value for Until_n_steps method
dt                  | {dt}              | float     | channel process timestep [sec]
code_file           | [site_prefix]_flow.rtg        | string    | grid of D8 flow codes
slope_file          | [site_prefix]_slope.rtg       | string    | grid of D8 slopes [m/m]
MANNING             | {manning}      | int       | option to use Manning's n for roughness
LAW_OF_WALL         | {low}      | int       | option to use Law of Wall for roughness
nval_type           | {nval_type}        | string    | allowed input types
nval                | {nval}             | {nval_kind}    | Manning's N values [s/m^(1/3)]
z0val_type          | {z0_type}       | string    | allowed input types
z0val               | {z0val}            | {z0_kind}    | Law-of-wall roughness values [m]
#===============================================================================
# Input 2
width_type          | {width_type}       | string    | allowed input types
width               | {width}            | {width_kind}    | bottom width of trapezoid [m]
angle_type          | {angle_type}       | string    | allowed input types
angle               | {angle}            | {angle_kind}    | bank angle of trapezoid [deg]
d0_type             | Scalar          | string    | allowed input types
d0                  | {d0}               | float     | initial flow depth [m]
sinu_type           | Scalar        | string    | allowed input types
sinu                | {sinu}             | float     | absolute channel sinuosity [m/m]
d_bankfull_type     | {dbank_type}          | string     | allowed input types
d_bankfull          | {dbank}    | {dbank_kind}      | channel bankfull water depth [m]
{flood_line}CHECK_STABILITY     | {check}     | string    | option to check depths and velocities every step
#===============================================================================
# Output 1
save_grid_dt        | 60.0     | float     | time interval between saved grids [sec]
SAVE_Q_GRIDS        | No     | string    | option to save computed Q grids {{Yes; No}}
Q_gs_file           | [case_prefix]_2D-Q.nc        | string    | filename for Q grid stack [m^3/s]
SAVE_U_GRIDS        | No     | string    | option to save computed u grids {{Yes; No}}
u_gs_file           | [case_prefix]_2D-u.nc        | string    | filename for u grid stack [m/s]
SAVE_D_GRIDS        | No     | string    | option to save computed d grids {{Yes; No}}
d_gs_file           | [case_prefix]_2D-d.nc        | string    | filename for d grid stack [m]
SAVE_F_GRIDS        | No     | string    | option to save computed f grids {{Yes; No}}
f_gs_file           | [case_prefix]_2D-f.nc        | string    | filename for f grid stack [none]
#===============================================================================
# Output 2
save_pixels_dt      | 60.0  | float      | time interval between time series values [sec]
pixel_file          | [case_prefix]_outlets.txt     | string    | filename for monitored pixel info
SAVE_Q_PIXELS       | No   | string     | option to save computed Q time series {{Yes; No}}
Q_ts_file           | [case_prefix]_0D-Q.txt        | string    | filename for computed Q time series
SAVE_U_PIXELS       | No    | string     | option to save computed u time series {{Yes; No}}
u_ts_file           | [case_prefix]_0D-u.txt        | string    | filename for computed u time series
SAVE_D_PIXELS       | No    | string     | option to save computed d time series {{Yes; No}}
d_ts_file           | [case_prefix]_0D-d.txt        | string    | filename for computed d time series
SAVE_F_PIXELS       | No    | string     | option to save computed f time series {{Yes; No}}
f_ts_file           | [case_prefix]_0D-f.txt        | string    | filename for computed f time series
"""

_D8_CFG = """#===============================================================================
# TopoFlow Config File for: D8_Global
#===============================================================================
comp_status         | Enabled       | string    | component status {{Enabled; Disabled}}
n_steps             | 1             | long      | number of time steps
dt                  | 1.0           | float     | process timestep [sec]
DEM_file            | [site_prefix]_DEM.rtg         | string    | filename of binary file with DEM
A_units             | km^2          | string    | area grid units, m^2 or km^2
LINK_FLATS          | {link_flats}      | long      | option to link flats, 0 or 1
FILL_PITS_IN_Z0     | 0      | long      | option to fill pits in original DEM, 0 or 1
LR_PERIODIC         | 0      | long      | B.C., periodic in left-right direction, 0 or 1
TB_PERIODIC         | 0      | long      | B.C., periodic in top-bottom direction, 0 or 1
save_grid_dt        | 1.0     | float     | time interval between saved grids [sec]
SAVE_AREA_GRIDS     | No  | string    | option to save grids of D8 area {{Yes; No}}
area_gs_file        | [case_prefix]_2D-A.nc     | string    | filename
SAVE_CODE_GRIDS     | No  | string    | option to save grids of D8 codes {{Yes; No}}
code_gs_file        | [case_prefix]_2D-flow.nc     | string    | filename
SAVE_DS_GRIDS       | No    | string    | option to save grids of ds {{Yes; No}}
ds_gs_file          | [case_prefix]_2D-ds.nc       | string    | filename
SAVE_DW_GRIDS       | No    | string    | option to save grids of dw {{Yes; No}}
dw_gs_file          | [case_prefix]_2D-dw.nc       | string    | filename
save_pixels_dt      | 1.0   | float     | time interval between time series values [sec]
pixel_file          | [case_prefix]_outlets.txt       | string    | filename for monitored pixel info
SAVE_AREA_PIXELS    | No | string    | option
area_ts_file        | [case_prefix]_0D-A.nc     | string    | filename
SAVE_CODE_PIXELS    | No | string    | option
code_ts_file        | [case_prefix]_0D-flow.nc     | string    | filename
SAVE_DS_PIXELS      | No   | string    | option
ds_ts_file          | [case_prefix]_0D-ds.nc       | string    | filename
SAVE_DW_PIXELS      | No   | string    | option
dw_ts_file          | [case_prefix]_0D-dw.nc       | string    | filename
"""


def write_case(case_dir, DEM, slope, width, angle, nval=None, z0val=None,
               site='Synth', case='Case1', dx=30.0, dy=30.0, dt=1.0,
               kinematic=True, manning=True, link_flats=1, outlet=None,
               n_steps=100, d0=0.0, sinu=1.0, scalar_angle=None,
               scalar_width=None, check_stability='Yes', byte_order=None, flood=False,
               d_bankfull=None, attenuate_days=None, geographic=None):
    """Write a stand-alone routing case into `case_dir` (all files side by
    side; in_directory = out_directory = '.').  Returns the channels CFG path."""
    os.makedirs(case_dir, exist_ok=True)
    ny, nx = DEM.shape
    finite = DEM[np.isfinite(DEM)]
    geo = {}
    if geographic is not None:
        # fixed-angle pixels: dx, dy in arcseconds, (south edge [deg N], west edge [deg E])
        geo = dict(pixel_geom=0, y_south_edge=geographic[0], x_west_edge=geographic[1])
    info = grid_io.make_rti(nx, ny, dx, dy, byte_order=byte_order,
                            grid_file=site + '_DEM.rtg',
                            gmin=float(finite.min()), gmax=float(finite.max()), **geo)
    p = lambda name: os.path.join(case_dir, name)
    grid_io.write_rti(p(site + '.rti'), info)
    grid_io.write_rtg(p(site + '_DEM.rtg'), DEM, info)
    grid_io.write_rtg(p(site + '_slope.rtg'), slope, info)
    if scalar_width is None:
        grid_io.write_rtg(p(site + '_chan-w.rtg'), width, info)
    if scalar_angle is None:
        grid_io.write_rtg(p(site + '_chan-a.rtg'), angle, info)
    if nval is not None and np.ndim(nval) == 2:
        grid_io.write_rtg(p(site + '_chan-n.rtg'), nval, info)
    if z0val is not None and np.ndim(z0val) == 2:
        grid_io.write_rtg(p(site + '_chan-z0.rtg'), z0val, info)
    if d_bankfull is not None and np.ndim(d_bankfull) == 2:
        grid_io.write_rtg(p(site + '_d-bank.rtg'), d_bankfull, info)
    with open(p(case + '_path_info.cfg'), 'w') as fh:
        fh.write('#' + '=' * 79 + '\n# TopoFlow Config File for: Path_Information\n'
                 '#' + '=' * 79 + '\n'
                 'in_directory   | .    | string  | input directory\n'
                 'out_directory  | .    | string  | output directory\n'
                 'site_prefix    | %s   | string  | file prefix for the study site\n'
                 'case_prefix    | %s   | string  | file prefix for the model scenario\n'
                 % (site, case))
    with open(p(case + '_time_info.cfg'), 'w') as fh:
        fh.write('#' + '=' * 79 + '\n# TopoFlow Config File for: Time_Information\n'
                 '#' + '=' * 79 + '\n'
                 'start_date   |  2020-01-01   | string  | start date for model run\n'
                 'start_time   |  00:00:00     | string  | start time (HH:MM:SS)\n'
                 'end_date     |  2020-01-02   | string  | end date for model run\n'
                 'end_time     |  00:00:00     | string  | end time (HH:MM:SS)\n')
    if outlet is None:
        outlet = (ny - 2, nx // 2)
    with open(p(case + '_outlets.txt'), 'w') as fh:
        fh.write('\n' + '-' * 60 + '\n Monitored Grid Cell (Outlet) Information\n'
                 + '-' * 60 + '\n    Column       Row     Area [km^2]      Relief [m]\n'
                 + '-' * 60 + '\n')
        fh.write('%10d %9d %15.6f %15.4f\n' % (outlet[1], outlet[0], 1.0, 1.0))
    with open(p(case + '_d8_global.cfg'), 'w') as fh:
        fh.write(_D8_CFG.format(link_flats=int(link_flats)))

    def gs(is_grid, fname, val):
        return (('Grid', '[site_prefix]' + fname, 'string') if is_grid
                else ('Scalar', repr(float(val)), 'float'))
    nv = gs(nval is not None and np.ndim(nval) == 2, '_chan-n.rtg',
            0.03 if nval is None else (nval if np.ndim(nval) == 0 else 0))
    z0 = gs(z0val is not None and np.ndim(z0val) == 2, '_chan-z0.rtg',
            0.01 if z0val is None else (z0val if np.ndim(z0val) == 0 else 0))
    wd = gs(scalar_width is None, '_chan-w.rtg', scalar_width or 0)
    an = gs(scalar_angle is None, '_chan-a.rtg', scalar_angle or 0)
    db = gs(d_bankfull is not None and np.ndim(d_bankfull) == 2, '_d-bank.rtg',
            10.0 if d_bankfull is None or np.ndim(d_bankfull) == 2 else d_bankfull)
    flood_line = ''
    if attenuate_days is not None:
        flood_line += ('ATTENUATE           | 1     | int        | route runoff through a linear reservoir\n'
                       'n_days              | %r    | float      | drainage time of the reservoir [days]\n'
                       % float(attenuate_days))
    if flood:
        flood_line += ('FLOOD_OPTION        | 1     | int        | option to model overbank flow\n'
                      'SAVE_DF_GRIDS       | No     | string    | option to save d_flood grids\n'
                      'd_flood_gs_file     | [case_prefix]_2D-d-flood.nc  | string    | filename\n'
                      'SAVE_DF_PIXELS      | No     | string    | option to save d_flood series\n'
                      'd_flood_ts_file     | [case_prefix]_0D-d-flood.txt | string    | filename\n')
    ext = '_channels_kinematic_wave.cfg' if kinematic else '_channels_diffusive_wave.cfg'
    cfg = p(case + ext)
    with open(cfg, 'w') as fh:
        fh.write(_CHAN_CFG.format(
            title='Kinematic_Wave' if kinematic else 'Diffusive_Wave',
            n_steps=int(n_steps), dt=repr(float(dt)),
            manning=1 if manning else 0, low=0 if manning else 1,
            nval_type=nv[0], nval=nv[1], nval_kind=nv[2],
            z0_type=z0[0], z0val=z0[1], z0_kind=z0[2],
            width_type=wd[0], width=wd[1], width_kind=wd[2],
            angle_type=an[0], angle=an[1], angle_kind=an[2],
            d0=repr(float(d0)), sinu=repr(float(sinu)), check=check_stability,
            dbank_type=db[0], dbank=db[1], dbank_kind=db[2], flood_line=flood_line))
    return cfg
