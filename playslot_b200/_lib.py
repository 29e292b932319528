""" ctypes loader of libplayslot_b200.so (the C ABI of include/playslot_b200.h).  Fails loudly. """
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libplayslot_b200.so")

c_f = ctypes.c_void_p      # device pointers travel as void*
c_i = ctypes.c_int
c_fl = ctypes.c_float

# name -> (restype, argtypes); must list every symbol declared in include/playslot_b200.h
SIGNATURES = {
    "ps_version": (c_i, []),
    "ps_last_error": (ctypes.c_char_p, []),
    "ps_launch_count": (ctypes.c_ulonglong, []),
    "ps_reset_launch_count": (None, []),
    "ps_device_info": (c_i, [ctypes.POINTER(c_i)] * 3),
    "ps_set_sm_reserve": (c_i, [c_i]),
    "ps_get_sm_reserve": (c_i, []),
    "ps_linear": (c_i, [c_f, c_i, c_f, c_f, c_f, c_f, c_fl, c_f, c_i, c_i, c_i, c_f, c_i, c_i, c_i, c_i, c_i, c_f]),
    "ps_linear_tc": (c_i, [c_f, c_i, c_f, c_fl, c_i, c_i, c_f, c_f, c_f, c_fl, c_f, c_i, c_i, c_i, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_f]),
    "ps_linear_tc_planes": (c_i, [c_f, c_i, c_f, ctypes.c_longlong, c_f, c_fl, c_i, c_i, c_f, c_f, c_f, c_fl, c_f, c_i, c_i, c_i, c_f, c_f,
                                   ctypes.c_longlong, c_i, c_i, c_i, c_i, c_i, c_f]),
    "ps_mha_planes": (c_i, [c_f, c_i, c_f, c_f, ctypes.c_longlong, c_i, c_i, c_i, c_i, c_i, c_i, c_f]),
    "ps_mha": (c_i, [c_f, c_i, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_f]),
    "ps_conv2d": (c_i, [c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f]),
    "ps_conv2d_split": (c_i, [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f]),
    "ps_conv2d_split_u8": (c_i, [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f]),
    "ps_sa_partial_floats": (c_i, [c_i]),
    "ps_sa_main": (c_i, [c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_fl, c_fl, c_f]),
    "ps_sa_main_tc": (c_i, [c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_fl, c_fl, c_f]),
    "ps_sa_prepare": (c_i, [c_f, c_f, ctypes.c_longlong, c_i, c_fl, c_f]),
    "ps_enc_mlp_planes": (c_i, [c_f, c_i, c_i, c_f, c_f, c_fl, c_f, c_fl, c_f, c_f, c_fl, c_f, c_fl, c_f, ctypes.c_longlong, c_f]),
    "ps_enc_mlp_planes_paired": (c_i, [c_f, c_i, c_i, c_f, c_f, c_fl, c_f, c_fl, c_f, c_f, c_fl, c_f, c_fl, c_f, ctypes.c_longlong, c_i, c_f]),
    "ps_sa_tc2_chunks": (c_i, [c_i, c_i]),
    "ps_sa_main_tc2": (c_i, [c_f, ctypes.c_longlong, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_fl, c_f]),
    "ps_sa_update": (c_i, [c_f, c_i, c_f, c_f] + [c_f] * 17 + [c_f, c_f, c_f] + [c_i, c_i, c_i, c_i, c_fl, c_fl, c_i, c_i, c_f]),
    "ps_sa_update2": (c_i, [c_f, c_i, c_f, c_f] + [c_f] * 16 + [c_f, c_f] + [c_i, c_i, c_i, c_i, c_fl, c_i, c_i, c_f]),
    "ps_dec_expand": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f]),
    "ps_dec_composite": (c_i, [c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f]),
    "ps_invdyn_tail": (c_i, [c_f] * 9 + [c_i, c_i, c_i, c_i, c_i, c_i, c_fl, c_f]),
    "ps_vq_lookup": (c_i, [c_f] * 6 + [c_i, c_i, c_i, c_f]),
    "ps_conv_tc_plane_elems": (ctypes.c_longlong, [c_i, c_i, c_i]),
    "ps_conv5x5_tc": (c_i, [c_f, c_f, c_f, c_fl, c_f, c_f, c_f] + [c_i] * 6 + [c_f]),
    "ps_dec_expand_split": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f]),
    "ps_layernorm_fwd": (c_i, [c_f, c_f, c_f, c_fl, c_f, c_f, c_f, c_i, c_i, c_f]),
    "ps_layernorm_bwd": (c_i, [c_f] * 8 + [c_i, c_i, c_f]),
    "ps_linear_wgrad": (c_i, [c_f, c_i, c_f, c_i, c_f, c_f, c_i, c_i, c_i, c_f]),
    "ps_linear_wgrad_acc": (c_i, [c_f, c_i, c_f, c_i, c_f, c_f, c_i, c_i, c_i, c_f]),
    "ps_act_fwd": (c_i, [c_f, c_f, ctypes.c_longlong, c_i, c_f]),
    "ps_act_bwd": (c_i, [c_f, c_f, c_f, ctypes.c_longlong, c_i, c_f]),
    "ps_mha_bwd": (c_i, [c_f, c_i, c_f, c_i, c_f, c_i, c_i, c_i, c_i, c_i, c_f]),
    "ps_invdyn_tail_bwd": (c_i, [c_f] * 11 + [c_i] * 6 + [c_fl, c_f]),
    "ps_dec_composite_bwd": (c_i, [c_f] * 6 + [c_i, c_i, c_i, c_i, c_f]),
    "ps_dec_final_dgrad": (c_i, [c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f]),
    "ps_dec_final_dgrad_planes": (c_i, [c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f]),
    "ps_conv5x5_tc_dgrad": (c_i, [c_f, c_f, c_fl, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f]),
    "ps_dec_expand_bwd": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f]),
    "ps_dec_composite_tc": (c_i, [c_f, c_f, c_fl, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f]),
    "ps_nchw_to_split": (c_i, [c_f, c_f, c_i, c_i, c_i, c_i, c_f]),
    "ps_sa_core_chunks": (c_i, [c_i]),
    "ps_sa_core_fwd": (c_i, [c_f] * 7 + [c_i, c_i, c_i, c_fl, c_f]),
    "ps_sa_core_bwd": (c_i, [c_f] * 8 + [c_i, c_i, c_i, c_fl, c_i, c_f]),
    "ps_gru_gates_fwd": (c_i, [c_f] * 5 + [ctypes.c_longlong, c_i, c_f]),
    "ps_gru_gates_bwd": (c_i, [c_f] * 7 + [ctypes.c_longlong, c_i, c_f]),
    "ps_conv5x5_planes_wgrad": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_fl, c_f]),
    "ps_conv5x5_planes_wgrad_tc": (c_i, [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_fl, c_f]),
    "ps_planes_colsum": (c_i, [c_f, c_f, c_i, c_i, c_i, c_i, c_fl, c_f]),
    "ps_conv_first_wgrad": (c_i, [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_fl, c_f]),
    "ps_conv3x3_final_wgrad": (c_i, [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_f]),
    "ps_conv3x3_final_wgrad_planes": (c_i, [c_f, c_f, c_i, c_f, c_f, c_i, c_i, c_i, c_f]),
    "ps_colsum": (c_i, [c_f, c_f, ctypes.c_longlong, ctypes.c_longlong, c_fl, c_f]),
    "ps_nhwc_to_split_scaled": (c_i, [c_f, c_i, c_f, c_f, c_i, c_i, c_i, c_i, c_f]),
    "ps_tblock_stack": (c_i, [c_f, c_i, c_i, c_i, c_i, c_i, c_fl, c_fl, c_fl] + [c_f] * 10 + [c_f, c_f]),
    "ps_tblock_stack_workspace_floats": (ctypes.c_longlong, [c_i, c_i]),
    "ps_tblock_stack_dbg": (c_i, [c_f, c_i, c_i, c_i, c_i, c_i, c_fl, c_fl, c_fl] + [c_f] * 10 + [c_f, c_f, c_f]),
    "ps_frame_sq_err": (c_i, [c_f, c_f, c_f, ctypes.c_longlong, c_i] + [ctypes.c_longlong] * 5 + [c_i, c_f]),
    "ps_mse_bwd": (c_i, [c_f, c_f, c_f, c_fl, c_f, ctypes.c_longlong, c_i, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, c_f]),
}

_lib = None


class PlaySlotKernelError(RuntimeError):
    pass


def load():
    """ Load the shared library once.  Raises if it has not been built: there is no fallback path. """
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PlaySlotKernelError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  playslot_b200 has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().ps_last_error().decode("utf-8", "replace")
        raise PlaySlotKernelError(f"{what} failed with status {status}: {msg}")
