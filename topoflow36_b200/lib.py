"""ctypes binding of libtfb_b200.so (C ABI declared in include/tfb.h).

There is NO fallback: if the shared library is missing or does not export a
symbol, importing fails loudly.  The library itself is only usable on a CUDA
device (every compute entry point returns TFB_ERR_CUDA otherwise).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = (os.environ.get('TFB_LIB_PATH') or os.path.join(_HERE, 'libtfb_b200.so'))  # override: kernel experiments

c_dbl_p = C.POINTER(C.c_double)
c_flt_p = C.POINTER(C.c_float)
vp = C.c_void_p

TFB_OK = 0
TFB_ERR_CUDA, TFB_ERR_ARG, TFB_ERR_NO_DEVICE, TFB_ERR_NOT_CONVERGED = -1, -2, -3, -4

CH_DIFFUSIVE = 1 << 0
CH_LAW_OF_WALL = 1 << 1
CH_DIAG = 1 << 2
CH_WRITE_R = 1 << 3
CH_ET_INPLACE = 1 << 4
CH_FLOOD = 1 << 5
CH_ATTENUATE = 1 << 6
CH_SRC_MET = 1 << 8
CH_SRC_SNOW = 1 << 9
CH_SRC_EVAP = 1 << 10
CH_SRC_INFIL = 1 << 11
CH_REVERSE = 1 << 12
NSUM = 24


class TfbError(RuntimeError):
    pass


class SG(C.Structure):
    """tfb_sg: scalar-or-grid fp64 input."""
    _fields_ = [('ptr', vp), ('val', C.c_double)]


class ChannelsScalars(C.Structure):
    """tfb_channels_scalars (device-resident 0-D values of the component)."""
    _fields_ = ([(n, C.c_double) for n in (
        'vol_R', 'vol_Q', 'vol_edge', 'Q_outlet', 'u_outlet', 'd_outlet',
        'f_outlet', 'Q_peak', 'T_peak', 'u_peak', 'Tu_peak', 'd_peak',
        'Td_peak', 'vol_P', 'vol_PR', 'vol_PS', 'vol_SM', 'vol_ET', 'vol_IN',
        'dt_cfl_min')] +
        [(n, C.c_int64) for n in (
            'n_neg_d', 'n_nan_d', 'n_inf_d', 'n_neg_u', 'n_nan_u', 'n_inf_u',
            'n_nan_IN', 'step_count', 'n_step_failed')])


class ChannelsStepArgs(C.Structure):
    """tfb_channels_step_args -- field order must match include/tfb.h."""
    _fields_ = [
        ('ny', C.c_int32), ('nx', C.c_int32),
        ('row0', C.c_int32), ('ny_global', C.c_int32),
        ('flags', C.c_uint32),
        ('outlet_row', C.c_int32), ('outlet_col', C.c_int32),
        ('is_R_grid', C.c_int32),
        ('dt', C.c_double), ('time_min', C.c_double), ('rho_H2O', C.c_double),
        ('n_pixels_global', C.c_int64),
        ('halo', C.c_int32), ('pitch', C.c_int32),
        ('geo', vp), ('dx', vp), ('dy', vp), ('dd', vp),
        ('sinu', SG), ('da', SG), ('S_bed', vp), ('sqrtS', vp), ('width', SG),
        ('tan_angle', SG), ('cos_angle', SG), ('rough', SG),
        ('vol', vp), ('vol_out', vp), ('Q_in', vp), ('Q_out', vp),
        ('d', vp), ('u', vp),
        ('Rh', vp), ('A_wet', vp), ('P_wet', vp), ('f', vp), ('tau', vp),
        ('u_star', vp), ('froude', vp), ('S_free', vp), ('R', vp),
        ('P_rain', SG), ('SM', SG), ('GW', SG), ('MR', SG), ('IN', SG),
        ('ET', SG), ('ET_rw', vp),
        ('T_air', SG), ('T_rain_snow', C.c_double),
        ('c0', SG), ('T0', SG), ('h_swe', vp), ('h_swe_out', vp),
        ('h_snow', vp), ('rho_snow', C.c_double),
        ('T_surf', SG), ('Qn_SW', SG), ('Qn_LW', SG),
        ('alpha', C.c_double), ('K_soil', C.c_double), ('soil_x', C.c_double),
        ('T_soil_x', C.c_double),
        ('Ks', SG), ('Ki', SG), ('qs', SG), ('qi', SG), ('G', SG),
        ('I', vp), ('I_out', vp), ('h_table', SG), ('elev', SG),
        ('scalars', vp), ('partials', vp), ('ticket', vp),
        ('finalize', C.c_int32),
        ('scalars_mirror', vp),
        ('vol_bankfull', SG), ('d_flood', vp), ('vol_flood', vp),
        ('flood_manning_n', C.c_double), ('width_ratio', C.c_double),
        ('vol_stored', vp), ('vol_stored_out', vp), ('t_drain', C.c_double),
    ]


class MetEnergyArgs(C.Structure):
    """tfb_met_energy_args."""
    _IN = ('T_air', 'RH', 'uz', 'z', 'z0_air', 'h_snow', 'h_ice', 'p0', 'albedo', 'em_surf',
           'lat_deg', 'alpha', 'beta', 'TSN_offset', 'cloud_factor', 'canopy_factor', 'Qa', 'Qc')
    _OUT = ('e_sat_air', 'e_air', 'T_dew', 'T_surf', 'e_sat_surf', 'Ri', 'Dn', 'Dh', 'Qh', 'W_p',
            'e_surf', 'Qe', 'Qn_SW', 'em_air', 'Qn_LW', 'Q_sum')
    _fields_ = ([('ny', C.c_int32), ('nx', C.c_int32), ('satterlund', C.c_int32)] +
                [(n, SG) for n in _IN] +
                [('julian_day', C.c_double), ('dust_atten', C.c_double)] +
                [(n, vp) for n in _OUT])


class PeerHalo(C.Structure):
    """tfb_peer_halo: neighbours' halo rows and step flags mapped with CUDA IPC."""
    _fields_ = [('up_Q', vp), ('down_Q', vp), ('up_flag', vp), ('down_flag', vp),
                ('from_up', vp), ('from_down', vp), ('step', C.c_uint32)]


class ChannelsPacked(C.Structure):
    """tfb_channels_packed: compact static inputs of the packed fast path."""
    _fields_ = [('geo32', vp), ('width32', vp), ('coef', vp), ('angle_lut', vp),
                ('row_geom', vp), ('outlet_rough', C.c_double), ('n_angles', C.c_int32),
                ('inv_rough', vp), ('S32', vp), ('noflow_idx', vp), ('n_noflow', C.c_int64),
                ('rough64', vp), ('lut_tc', vp), ('width64', vp), ('S64', vp), ('sinu64', vp)]


_i32, _i64, _f32, _f64 = C.c_int32, C.c_int64, C.c_float, C.c_double

# name -> (restype, argtypes); every symbol include/tfb.h declares
SIGNATURES = {
    'tfb_abi_version': (C.c_int, []),
    'tfb_last_error': (C.c_char_p, []),
    'tfb_stream_synchronize': (C.c_int, [vp]),
    'tfb_device_info': (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'tfb_d8_resolve_table': (C.c_int, [vp]),
    'tfb_d8_codes': (C.c_int, [vp, _i32, _i32, _i32, _i32, _i32, _f32, _f32,
                               _f32, _f64, _i32, vp, vp]),
    'tfb_d8_break_ties': (C.c_int, [vp, _i64, vp]),
    'tfb_d8_link_flats_sweep': (C.c_int, [vp, vp, _i32, _i32, _i32, _i32, _i32,
                                          vp, vp, vp]),
    'tfb_d8_link_flats': (C.c_int, [vp, vp, _i32, _i32, _i32,
                                    C.POINTER(_i32), vp]),
    'tfb_d8_finalize_codes': (C.c_int, [vp, vp, _i64, vp]),
    'tfb_d8_ds_dw': (C.c_int, [vp, vp, vp, vp, _i32, _i32, vp, vp, vp]),
    'tfb_d8_area': (C.c_int, [vp, _i32, _i32, _f64, vp, vp, vp, vp,
                              C.POINTER(_i64), vp]),
    'tfb_d8_area_slab_init': (C.c_int, [vp, _i32, _i32, _i32, _i32, vp, vp, vp]),
    'tfb_d8_area_slab_walk': (C.c_int, [_i32, _i32, _f64, vp, vp, vp, _i32, vp]),
    'tfb_d8_area_unpack': (C.c_int, [vp, _i64, vp, vp, C.POINTER(_i64), vp]),
    'tfb_d8_slope': (C.c_int, [vp, vp, vp, vp, vp, _i32, _i32, vp, vp]),
    'tfb_d8_parent_ids': (C.c_int, [vp, _i32, _i32, vp, vp]),
    'tfb_d8_aspect': (C.c_int, [vp, _i64, vp, vp]),
    'tfb_d8_new_slope': (C.c_int, [vp, vp, vp, vp, vp, _i32, _i32, _f32, vp, vp,
                                   C.POINTER(_i32), vp]),
    'tfb_fill_pits_scratch_bytes': (_i64, [_i32, _i32]),
    'tfb_fill_pits': (C.c_int, [vp, _i32, _i32, _f32, _f32, _i32, vp, C.POINTER(_i64),
                                C.POINTER(_i32), vp]),
    'tfb_area_range': (C.c_int, [vp, _i64, C.POINTER(_f32), C.POINTER(_f32), vp]),
    'tfb_power_law_grid': (C.c_int, [vp, _i64, _f64, _f64, _i32, vp, vp, vp]),
    'tfb_sizeof_channels_step_args': (_i64, []),
    'tfb_sizeof_channels_scalars': (_i64, []),
    'tfb_channels_partials_len': (_i64, [_i32, _i32]),
    'tfb_channels_pack_geo': (C.c_int, [vp, _i32, _i32, _i32, vp, vp]),
    'tfb_channels_step': (C.c_int, [C.POINTER(ChannelsStepArgs), vp]),
    'tfb_sizeof_channels_packed': (_i64, []),
    'tfb_channels_pack_geo32': (C.c_int, [vp, vp, _i32, _i32, _i32, vp, vp]),
    'tfb_channels_pack_rows': (C.c_int, [vp, vp, vp, _f64, vp, _f64, _i32, vp, vp]),
    'tfb_channels_step_packed': (C.c_int, [C.POINTER(ChannelsStepArgs),
                                           C.POINTER(ChannelsPacked), vp]),
    'tfb_channels_step_packed_dw_a': (C.c_int, [C.POINTER(ChannelsStepArgs),
                                                C.POINTER(ChannelsPacked), vp]),
    'tfb_channels_step_ext': (C.c_int, [C.POINTER(ChannelsStepArgs), C.POINTER(ChannelsPacked), vp]),
    'tfb_channels_step_ext_dw_a': (C.c_int, [C.POINTER(ChannelsStepArgs),
                                             C.POINTER(ChannelsPacked), vp]),
    'tfb_channels_step_ext_dw_b': (C.c_int, [C.POINTER(ChannelsStepArgs),
                                             C.POINTER(ChannelsPacked), vp]),
    'tfb_channels_step_packed_dw_b': (C.c_int, [C.POINTER(ChannelsStepArgs),
                                                C.POINTER(ChannelsPacked), vp]),
    'tfb_sizeof_peer_halo': (_i64, []),
    'tfb_peer_alloc': (C.c_int, [_i64, C.POINTER(vp), C.c_char_p]),
    'tfb_peer_open': (C.c_int, [C.c_char_p, C.POINTER(vp)]),
    'tfb_peer_close': (C.c_int, [vp]),
    'tfb_peer_free': (C.c_int, [vp]),
    'tfb_channels_step_packed_p2p': (C.c_int, [C.POINTER(ChannelsStepArgs),
                                               C.POINTER(ChannelsPacked), C.POINTER(PeerHalo), vp]),
    'tfb_channels_step_packed_dw_a_p2p': (C.c_int, [C.POINTER(ChannelsStepArgs),
                                                    C.POINTER(ChannelsPacked), C.POINTER(PeerHalo), vp]),
    'tfb_channels_step_packed_dw_b_p2p': (C.c_int, [C.POINTER(ChannelsStepArgs),
                                                    C.POINTER(ChannelsPacked), C.POINTER(PeerHalo), vp]),
    'tfb_channels_step_packed_dw_c': (C.c_int, [C.POINTER(ChannelsStepArgs),
                                                C.POINTER(ChannelsPacked), vp]),
    'tfb_channels_fold': (C.c_int, [C.POINTER(ChannelsStepArgs), vp, vp]),
    'tfb_sources_partials_len': (_i64, []),
    'tfb_channels_minmax': (C.c_int, [vp, vp, vp, vp, _i32, _i32, vp, vp, vp,
                                      vp, vp]),
    'tfb_met_precip': (C.c_int, [SG, SG, _f64, SG, _f64, _i32, _i32, vp, vp,
                                 vp, vp, vp, vp]),
    'tfb_infil_green_ampt': (C.c_int, [SG, SG, SG, SG, SG, SG, SG, SG, SG, _i32,
                                       SG, _f64, _i32, _i32, vp, vp, vp, vp, vp,
                                       vp]),
    'tfb_snow_degree_day': (C.c_int, [SG, SG, SG, SG, _f64, _f64, SG, _f64,
                                      _i32, _i32, vp, vp, vp, vp, vp, vp, vp]),
    'tfb_met_albedo': (C.c_int, [SG, SG, SG, SG, _i32, _i32, _f64, _f64, _f64, _i32, vp, _i32,
                                 _i32, vp, vp, vp]),
    'tfb_sizeof_met_energy_args': (_i64, []),
    'tfb_met_energy': (C.c_int, [C.POINTER(MetEnergyArgs), vp]),
    'tfb_evap_priestley_taylor': (C.c_int, [SG, SG, SG, SG, _f64, _f64, _f64,
                                            _f64, SG, _f64, _i32, _i32, vp, vp,
                                            vp, vp, vp, vp]),
}

_lib = None


def load():
    """dlopen the library once, bind every declared symbol, check the ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TfbError(
            'libtfb_b200.so not found at %s -- build it with '
            '`python -c "import __graft_entry__ as g; g.build()"` or '
            '`bash topoflow36_b200/csrc/build.sh`. There is no CPU fallback.'
            % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise TfbError('libtfb_b200.so does not export %s' % name)
        fn.restype = res
        fn.argtypes = args
    if lib.tfb_abi_version() != 1:
        raise TfbError('libtfb_b200.so ABI version %d != 1' % lib.tfb_abi_version())
    if lib.tfb_sizeof_channels_step_args() != C.sizeof(ChannelsStepArgs):
        raise TfbError('tfb_channels_step_args layout mismatch: C %d vs ctypes %d'
                       % (lib.tfb_sizeof_channels_step_args(),
                          C.sizeof(ChannelsStepArgs)))
    if lib.tfb_sizeof_channels_scalars() != C.sizeof(ChannelsScalars):
        raise TfbError('tfb_channels_scalars layout mismatch')
    if lib.tfb_sizeof_peer_halo() != C.sizeof(PeerHalo):
        raise TfbError('tfb_peer_halo layout mismatch')
    if lib.tfb_sizeof_met_energy_args() != C.sizeof(MetEnergyArgs):
        raise TfbError('tfb_met_energy_args layout mismatch')
    if lib.tfb_sizeof_channels_packed() != C.sizeof(ChannelsPacked):
        raise TfbError('tfb_channels_packed layout mismatch')
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != TFB_OK:
        msg = load().tfb_last_error()
        raise TfbError('%s failed (%d): %s'
                       % (what or 'tfb call', rc, msg.decode() if msg else ''))


def sg(x):
    """Build a tfb_sg from a python float / 0-D value or a CUDA torch tensor."""
    import torch
    if isinstance(x, torch.Tensor):
        if x.dim() == 0:
            return SG(None, float(x))
        if not x.is_cuda or x.dtype != torch.float64 or not x.is_contiguous():
            raise TfbError('grid inputs must be contiguous CUDA float64 tensors')
        return SG(x.data_ptr(), 0.0)
    return SG(None, float(x))


def ptr(t):
    return None if t is None else t.data_ptr()


_raw_stream = None


def stream_ptr(device_index=None):
    """cudaStream_t of torch's CURRENT stream (of `device_index`, default: the current
    device).  With a device index this is one C call (~0.3 us); building the
    torch.cuda.Stream object costs ~5 us, which matters on a step that launches once."""
    global _raw_stream
    import torch
    if device_index is not None:
        if _raw_stream is None:
            _raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', False)
        if _raw_stream:
            return _raw_stream(device_index)
    return torch.cuda.current_stream(device_index).cuda_stream
