"""ctypes binding of libeve_b200.so (the C ABI declared in include/eve_b200.h).

This is the whole Python <-> CUDA boundary: every call passes raw device pointers taken from
torch tensors plus the current CUDA stream.  There is no fallback: if the shared library is
missing, or a call returns a non-zero code, a RuntimeError is raised.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libeve_b200.so')

c_float_p = C.c_void_p
_lib = None


class ConvParams(C.Structure):
    _fields_ = [(k, C.c_int) for k in ('n', 'h', 'w', 'cin', 'cout', 'ksize', 'stride', 'pad')]


class EyeNetCnnParams(C.Structure):
    _fields_ = [(k, C.c_int) for k in ('n', 'nf', 'h', 'w')]


class EyeNetTailParams(C.Structure):
    _fields_ = [(k, C.c_int) for k in ('batch', 'steps', 'nf', 'use_head_pose', 'rnn_type',
                                       'rnn_cells')]


class RefineNetParams(C.Structure):
    _fields_ = [(k, C.c_int) for k in ('batch', 'steps', 'in_channels', 'use_skip', 'rnn_type',
                                       'rnn_cells', 'nf')]


class HeatmapParams(C.Structure):
    _fields_ = [('n', C.c_int), ('hm_w', C.c_int), ('hm_h', C.c_int), ('screen_w', C.c_float),
                ('screen_h', C.c_float), ('sigma', C.c_float)]


class LabelArgs(C.Structure):
    _fields_ = [('n', C.c_int)] + [(k, C.c_void_p) for k in (
        'left_pog_px', 'right_pog_px', 'left_valid', 'right_valid', 'mm_per_px', 'left_o', 'right_o',
        'left_R', 'cam', 'left_pog_cm', 'right_pog_cm', 'o', 'pog_px', 'pog_cm', 'valid', 'g')]


LOSS_OPS = {'angular': 0, 'mse': 1, 'l1': 2, 'euclidean': 3, 'identity': 4}
LOSS_MAX_TERMS = 40


class LossTerm(C.Structure):
    _fields_ = [('op', C.c_int), ('dim', C.c_int), ('pred', C.c_void_p), ('gt', C.c_void_p),
                ('valid', C.c_void_p), ('valid2', C.c_void_p), ('dpred', C.c_void_p)]


class AdamParams(C.Structure):
    _fields_ = [('count', C.c_longlong), ('lr', C.c_float), ('beta1', C.c_float),
                ('beta2', C.c_float), ('eps', C.c_float), ('weight_decay', C.c_float),
                ('max_norm', C.c_float), ('grad_scale', C.c_float), ('step', C.c_int),
                ('step_dev', C.c_void_p), ('lr_dev', C.c_void_p)]


EYE_RNN_TYPES = {None: 0, 'RNN': 1, 'LSTM': 2, 'GRU': 3}
REFINE_RNN_TYPES = {None: 0, 'CRNN': 1, 'CLSTM': 2, 'CGRU': 3}

# name -> (restype, argtypes); kept in one place so tests can check every symbol of the header
_P, _I, _F, _Z = C.c_void_p, C.c_int, C.c_float, C.c_size_t
SIGNATURES = {
    'eve_version': (_I, []),
    'eve_last_error': (C.c_char_p, []),
    'eve_launch_count': (C.c_longlong, []),
    'eve_profile_enable': (None, [_I]),
    'eve_profile_reset': (None, []),
    'eve_profile_read': (_I, [_I, _P, _P, _P, _P]),
    'eve_profile_dump': (C.c_longlong, [_P, C.c_longlong]),
    'eve_probe_mma_rate': (_I, [_I, _I, _I, _I, _I, _I, _P, _P]),
    'eve_probe_mma_rate_swizzle': (_I, [_I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'eve_probe_mma_rate_issuers': (_I, [_I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'eve_probe_mma_rate_pair': (_I, [_I, _I, _I, _I, _P, _P]),
    'eve_set_conv_mode': (None, [_I]),
    'eve_get_conv_mode': (_I, []),
    'eve_set_option': (_I, [C.c_char_p, _I]),
    'eve_get_option': (_I, [C.c_char_p, _P]),
    'eve_conv2d_workspace_bytes': (_Z, [_P]),
    'eve_conv2d_describe': (_I, [_P, _P, _Z]),
    'eve_conv2d_fwd': (_I, [_P, _P, _P, _P, _P, _P, _Z, _P]),
    'eve_conv2d_dgrad': (_I, [_P, _P, _P, _P, _P, _Z, _P]),
    'eve_conv2d_wgrad': (_I, [_P, _P, _P, _P, _P, _P, _Z, _P]),
    'eve_instnorm_act_fwd': (_I, [_P, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P]),
    'eve_instnorm_act_bwd': (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _P, _Z, _P]),
    'eve_instnorm_fused_workspace_bytes': (_Z, [_I, _I, _I]),
    'eve_instnorm_fused_fwd': (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P,
                                    _P, _P, _P, _P, _P, _P]),
    'eve_instnorm_fused_bwd': (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P, _P,
                                    _P, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    'eve_in_relu_maxpool_fwd': (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    'eve_in_relu_maxpool_bwd': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    'eve_adaptive_maxpool_fwd': (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P]),
    'eve_adaptive_maxpool_bwd': (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    'eve_upsample_bilinear_fwd': (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P]),
    'eve_upsample_bilinear_bwd': (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P]),
    'eve_nchw_to_nhwc': (_I, [_P, _I, _I, _I, _I, _P, _P]),
    'eve_nhwc_to_nchw': (_I, [_P, _I, _I, _I, _I, _P, _P]),
    'eve_eyenet_cnn_saved_bytes': (_Z, [_P]),
    'eve_eyenet_cnn_workspace_bytes': (_Z, [_P]),
    'eve_eyenet_cnn_fwd': (_I, [_P, _P, _P, _P, _P, _Z, _P, _Z, _P]),
    'eve_eyenet_cnn_bwd': (_I, [_P, _P, _P, _P, _I, _P, _Z, _P, _Z, _P]),
    'eve_eyenet_tail_num_weights': (_I, [_P]),
    'eve_eyenet_tail_saved_bytes': (_Z, [_P]),
    'eve_eyenet_tail_workspace_bytes': (_Z, [_P]),
    'eve_eyenet_tail_fwd': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P, _Z, _P]),
    'eve_eyenet_tail_bwd': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _Z, _P, _Z, _P]),
    'eve_refinenet_num_weights': (_I, [_P]),
    'eve_refinenet_weight_name': (C.c_char_p, [_P, _I]),
    'eve_refinenet_saved_bytes': (_Z, [_P]),
    'eve_refinenet_workspace_bytes': (_Z, [_P]),
    'eve_refinenet_fwd': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P, _Z, _P]),
    'eve_refinenet_bwd': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _Z, _P, _Z, _P]),
    'eve_heatmap_fwd': (_I, [_P, _P, _P, _P]),
    'eve_heatmap_bwd': (_I, [_P, _P, _P, _P, _P]),
    'eve_soft_argmax_fwd': (_I, [_P, _P, _P, _P]),
    'eve_soft_argmax_bwd': (_I, [_P, _P, _P, _P, _P]),
    'eve_pog_fwd': (_I, [_I, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P]),
    'eve_pog_bwd': (_I, [_I, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P, _P]),
    'eve_combined_gaze_fwd': (_I, [_I, _P, _P, _P, _P, _P, _P]),
    'eve_combined_gaze_bwd': (_I, [_I, _P, _P, _P, _P, _P, _P, _P]),
    'eve_offset_augmentation_fwd': (_I, [_I, _I, _P, _P, _P, _I, _P, _P]),
    'eve_offset_augmentation_bwd': (_I, [_I, _I, _P, _P, _P, _I, _P, _P, _P]),
    'eve_labels_fwd': (_I, [_P, _P]),
    'eve_heatmap_labels_fwd': (_I, [_P, _P, _P, _I, _P, _P, _P]),
    'eve_gaze_history_scratch_bytes': (_Z, [_I, _I]),
    'eve_gaze_history_fwd': (_I, [_I, _I, _I, _P, _P, _F, _P, _P, _P, _Z, _P]),
    'eve_gaze_history_bwd': (_I, [_I, _I, _I, _P, _P, _F, _P, _P, _P, _Z, _P]),
    'eve_masked_losses_fwd': (_I, [_I, _P, _I, _I, _P, _P]),
    'eve_masked_losses_bwd': (_I, [_I, _P, _I, _I, _P, _P]),
    'eve_heatmap_frame_losses_fwd': (_I, [_I, _I, _P, _P, _P, _P, _P]),
    'eve_heatmap_frame_losses_bwd': (_I, [_I, _I, _P, _P, _P, _P, _P, _P]),
    'eve_preprocess_frames': (_I, [_P, _I, _I, _I, _I, _I, _I, _F, _F, _P, _I, _P, _P]),
    'eve_adam_clip_workspace_bytes': (_Z, [_P]),
    'eve_adam_clip_step': (_I, [_P, _P, _P, _P, _P, _P, _P, _Z, _P]),
}


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'libeve_b200.so is missing (%s): build it with `python -c "import __graft_entry__ as g; '
            'g.build()"` or `make -C eve_b200/csrc`; there is no fallback path.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().eve_last_error().decode('utf-8', 'replace')


def check(rc, what):
    if rc != 0:
        raise RuntimeError('%s failed (code %d): %s' % (what, rc, last_error()))


def set_option(name, value):
    """Process-wide tuning switch of the tensor-core path (include/eve_b200.h: eve_set_option)."""
    check(load().eve_set_option(name.encode(), int(value)), 'eve_set_option(%s)' % name)


def get_option(name):
    v = C.c_int(0)
    check(load().eve_get_option(name.encode(), C.byref(v)), 'eve_get_option(%s)' % name)
    return v.value


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be contiguous fp32/int32."""
    if t is None:
        return None
    assert t.is_contiguous(), 'eve_b200: tensor must be contiguous'
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def ptr_table(tensors):
    """Host array of device pointers (NULL for None)."""
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else ptr(t)
    return arr


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError('%s: the B200 path needs CUDA tensors (got %s); there is no CPU fallback'
                           % (what, t.device))


_workspaces = {}
_pins = 0          # captured CUDA graphs that have workspace pointers baked into their kernels
_retired = []      # outgrown buffers kept alive while any such graph exists


def workspace(nbytes, device, tag='ws'):
    """A per-(device, tag) scratch buffer that only ever grows (stream-ordered reuse).

    While a captured graph is registered (pin_workspaces) an outgrown buffer is NOT freed: the
    graph's kernels keep writing into it on replay, so handing its memory back to the caching
    allocator would corrupt whatever tensor receives it next."""
    key = (str(device), tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None and _pins > 0:
            _retired.append(buf)
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def pin_workspaces():
    """Called when a CUDA graph that uses the shared scratch buffers has been captured."""
    global _pins
    _pins += 1


def unpin_workspaces():
    global _pins
    _pins = max(_pins - 1, 0)
    if _pins == 0:
        del _retired[:]
