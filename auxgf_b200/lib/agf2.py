"""ctypes loader of libagf2_b200.so -- the B200 counterpart of ``auxgf/lib/agf2.py``.

Mirrors the reference shim (``auxgf/lib/agf2.py:63-145``): same function names, same argument
meaning (``gf_occ``/``gf_vir`` are ``Aux``-like objects exposing ``naux``, ``nphys``, ``e``), same
``vv``/``vev`` accumulate-and-return behaviour, and -- like the reference (``agf2.py:18-21``) --
a missing shared library is a ``RuntimeError`` at import time.  There is no CPU fallback: the
compute entry points raise ``Agf2B200Error`` when no sm_100 device is present.

Extra (not in the reference): ``os_factor``/``ss_factor`` keywords (SCS, ``auxgf/aux/dfrmp2.py:82-84``)
and the device-resident calls used by ``auxgf_b200.agf2`` so that the DF tensors stay in HBM.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.realpath(__file__))
LIB_PATH = os.path.join(_HERE, "libagf2_b200.so")

_f64 = np.float64
_in2 = np.ctypeslib.ndpointer(dtype=_f64, ndim=2, flags="C,A")
_in1 = np.ctypeslib.ndpointer(dtype=_f64, ndim=1, flags="C,A")
_out2 = np.ctypeslib.ndpointer(dtype=_f64, ndim=2, flags="C,A,W")
_i64 = ctypes.c_int64
_dbl = ctypes.c_double
_vp = ctypes.c_void_p

# every symbol include/agf2_b200.h declares (tests check they are all exported)
SYMBOLS = [
    "build_part_loop_rhf", "build_part_loop_uhf", "agf2_b200_last_error", "agf2_b200_version",
    "agf2_b200_set_device", "agf2_b200_set_workspace_bytes", "agf2_b200_build_part_loop_rhf",
    "agf2_b200_build_part_loop_uhf", "agf2_b200_moments_rhf_dev", "agf2_b200_moments_uhf_dev",
    "agf2_b200_compress", "agf2_b200_eigh", "agf2_b200_launch_count", "agf2_b200_last_kernel_ms",
    "agf2_b200_set_simple_kernels", "agf2_b200_set_overlap", "agf2_b200_last_stage_ms",
    "agf2_b200_set_coulomb_factorization", "agf2_b200_last_plan", "agf2_b200_build_part_loop_rhf_sharded",
    "agf2_b200_build_part_loop_uhf_sharded", "agf2_b200_plan_stats",
    # contexts, options, communicator (round 2)
    "agf2_b200_ctx_create", "agf2_b200_ctx_destroy", "agf2_b200_ctx_get_device", "agf2_b200_ctx_synchronize",
    "agf2_b200_ctx_set_option", "agf2_b200_set_deterministic", "agf2_b200_sticky_status",
    "agf2_b200_ctx_last_stage_ms", "agf2_b200_ctx_last_plan", "agf2_b200_ctx_last_comm_bytes",
    "agf2_b200_comm_get_unique_id", "agf2_b200_comm_init", "agf2_b200_comm_info", "agf2_b200_comm_allreduce",
    "agf2_b200_comm_allgather", "agf2_b200_comm_broadcast", "agf2_b200_comm_barrier", "agf2_b200_comm_destroy",
    "agf2_b200_ctx_moments_rhf_dev", "agf2_b200_ctx_moments_uhf_dev", "agf2_b200_ctx_build_part_loop_rhf",
    "agf2_b200_ctx_build_part_loop_uhf", "agf2_b200_host_register", "agf2_b200_host_unregister", "agf2_b200_ctx_gemm",
    "agf2_b200_ctx_flop_counters",
    # pole-sharded builds
    "agf2_b200_pole_block", "agf2_b200_ctx_moments_rhf_dev_sharded", "agf2_b200_ctx_last_p2p_bytes",
    "agf2_b200_shard_cover", "agf2_b200_ctx_shard_sub_block",
]


class Agf2B200Error(RuntimeError):
    pass


def _load():
    try:
        clib = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        print(e)
        raise RuntimeError("Cannot locate agf2_b200 CUDA library (%s); build it with "
                           "`python -c 'import __graft_entry__ as g; g.build()'` or "
                           "`make -C auxgf_b200/csrc`." % LIB_PATH)
    return clib


_libagf2 = _load()

_libagf2.agf2_b200_last_error.restype = ctypes.c_char_p
_libagf2.agf2_b200_solver_last_error.restype = ctypes.c_char_p
_libagf2.agf2_b200_version.restype = ctypes.c_char_p
_libagf2.agf2_b200_launch_count.restype = _i64
_libagf2.agf2_b200_launch_count.argtypes = [ctypes.c_int]
_libagf2.agf2_b200_last_kernel_ms.argtypes = [ctypes.POINTER(_dbl), ctypes.POINTER(_dbl)]
_libagf2.agf2_b200_set_device.argtypes = [ctypes.c_int]
_libagf2.agf2_b200_set_workspace_bytes.argtypes = [_i64]
_libagf2.agf2_b200_set_simple_kernels.argtypes = [ctypes.c_int]
_libagf2.agf2_b200_plan_stats.argtypes = [_i64] * 8 + [_dbl, _dbl, ctypes.c_int, _i64, _i64, ctypes.POINTER(_i64),
                                          ctypes.POINTER(_i64)]
_libagf2.agf2_b200_build_part_loop_rhf.argtypes = [_in2, _in2, _in1, _in1] + [_i64] * 6 + [_dbl, _dbl, _out2, _out2]
_libagf2.agf2_b200_build_part_loop_rhf_sharded.argtypes = [_in2, _in2, _in1, _in1] + [_i64] * 6 + [_dbl, _dbl, _i64, _i64,
                                                                                                _out2, _out2]
_libagf2.agf2_b200_build_part_loop_uhf_sharded.argtypes = [_in2, _in2, _in2, _in1, _in1, _in1, _in1] + [_i64] * 8 + \
    [_dbl, _dbl, _i64, _i64, _out2, _out2]
_libagf2.agf2_b200_build_part_loop_uhf.argtypes = [_in2, _in2, _in2, _in1, _in1, _in1, _in1] + [_i64] * 8 + \
    [_dbl, _dbl, _out2, _out2]
_libagf2.agf2_b200_moments_rhf_dev.argtypes = [_vp, _i64, _vp, _i64, _vp, _vp] + [_i64] * 6 + [_dbl, _dbl, _i64, _i64,
                                                                                             _vp, _vp, _vp, ctypes.c_int]
_libagf2.agf2_b200_moments_uhf_dev.argtypes = [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp] + [_i64] * 8 + \
    [_dbl, _dbl, _i64, _i64, _vp, _vp, _vp, ctypes.c_int]
_libagf2.agf2_b200_compress.argtypes = [_vp, _vp, _i64, _vp, _vp, ctypes.c_int]
_libagf2.agf2_b200_eigh.argtypes = [_vp, _i64, _vp, _vp, ctypes.c_int]
_libagf2.agf2_b200_ctx_create.argtypes = [ctypes.c_int, ctypes.POINTER(_vp)]
_libagf2.agf2_b200_ctx_destroy.argtypes = [_vp]
_libagf2.agf2_b200_ctx_get_device.argtypes = [_vp, ctypes.POINTER(ctypes.c_int)]
_libagf2.agf2_b200_ctx_synchronize.argtypes = [_vp]
_libagf2.agf2_b200_ctx_set_option.argtypes = [_vp, ctypes.c_char_p, _i64]
_libagf2.agf2_b200_set_deterministic.argtypes = [ctypes.c_int]
_libagf2.agf2_b200_sticky_status.argtypes = [ctypes.c_int]
_libagf2.agf2_b200_ctx_last_stage_ms.argtypes = [_vp, ctypes.POINTER(_dbl)]
_libagf2.agf2_b200_ctx_last_plan.argtypes = [_vp, ctypes.POINTER(_i64)]
_libagf2.agf2_b200_ctx_last_comm_bytes.argtypes = [_vp, ctypes.POINTER(_i64)]
_libagf2.agf2_b200_comm_get_unique_id.argtypes = [_vp]
_libagf2.agf2_b200_comm_init.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_int]
_libagf2.agf2_b200_comm_info.argtypes = [_vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
_libagf2.agf2_b200_comm_allreduce.argtypes = [_vp, _vp, _i64, ctypes.c_int, _vp]
_libagf2.agf2_b200_comm_allgather.argtypes = [_vp, _vp, _i64, _vp]
_libagf2.agf2_b200_comm_broadcast.argtypes = [_vp, _vp, _i64, ctypes.c_int, _vp]
_libagf2.agf2_b200_comm_barrier.argtypes = [_vp]
_libagf2.agf2_b200_comm_destroy.argtypes = [_vp]
_libagf2.agf2_b200_ctx_moments_rhf_dev.argtypes = [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp] + [_i64] * 6 + \
    [_dbl, _dbl, _vp, _vp, _vp, ctypes.c_int]
_libagf2.agf2_b200_ctx_moments_uhf_dev.argtypes = [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp] + \
    [_i64] * 8 + [_dbl, _dbl, _vp, _vp, _vp, ctypes.c_int]
_libagf2.agf2_b200_ctx_build_part_loop_rhf.argtypes = [_vp, _in2, _in2, _in1, _in1] + [_i64] * 6 + [_dbl, _dbl, _out2, _out2]
_libagf2.agf2_b200_ctx_build_part_loop_uhf.argtypes = [_vp, _in2, _in2, _in2, _in1, _in1, _in1, _in1] + [_i64] * 8 + \
    [_dbl, _dbl, _out2, _out2]
_libagf2.agf2_b200_ctx_gemm.argtypes = [_vp, _vp, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _i64, ctypes.c_int, ctypes.c_int, _vp]
_libagf2.agf2_b200_ctx_flop_counters.argtypes = [_vp, ctypes.POINTER(_dbl), ctypes.c_int]
_libagf2.agf2_b200_pole_block.argtypes = [_i64, _i64, _i64, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]
_libagf2.agf2_b200_ctx_moments_rhf_dev_sharded.argtypes = [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp] + [_i64] * 4 + \
    [_dbl, _dbl, _vp, _vp, _vp, ctypes.c_int]
_libagf2.agf2_b200_ctx_last_p2p_bytes.restype = _i64
_libagf2.agf2_b200_ctx_last_p2p_bytes.argtypes = [_vp]
_libagf2.agf2_b200_shard_cover.argtypes = [_i64, _i64, _i64, _i64, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]
_libagf2.agf2_b200_ctx_shard_sub_block.restype = _i64
_libagf2.agf2_b200_ctx_shard_sub_block.argtypes = [_vp, _i64, _i64, _i64, _i64]
_libagf2.agf2_b200_host_register.argtypes = [_vp, _i64]
_libagf2.agf2_b200_host_unregister.argtypes = [_vp]
# the reference's own void symbols (drop-in ABI; argtypes as auxgf/lib/agf2.py:28-60)
_u32 = ctypes.c_uint32
_libagf2.build_part_loop_rhf.restype = None
_libagf2.build_part_loop_rhf.argtypes = [_in2, _in2, _in1, _in1] + [_u32] * 6 + [_out2, _out2]
_libagf2.build_part_loop_uhf.restype = None
_libagf2.build_part_loop_uhf.argtypes = [_in2, _in2, _in2, _in1, _in1, _in1, _in1] + [_u32] * 8 + [_out2, _out2]


def check(status):
    if status != 0:
        msg = _libagf2.agf2_b200_last_error().decode() or _libagf2.agf2_b200_solver_last_error().decode()
        raise Agf2B200Error("libagf2_b200 error %d: %s" % (status, msg))


def version():
    return _libagf2.agf2_b200_version().decode()


def launch_count(reset=False):
    return int(_libagf2.agf2_b200_launch_count(1 if reset else 0))


def last_kernel_ms():
    a, b = _dbl(0), _dbl(0)
    _libagf2.agf2_b200_last_kernel_ms(ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def set_simple_kernels(on):
    check(_libagf2.agf2_b200_set_simple_kernels(1 if on else 0))


def last_stage_ms():
    """Device ms of the last synchronous moments call: (k1_form, k2_moment pairs, M build, Coulomb step)."""
    out = (_dbl * 4)()
    check(_libagf2.agf2_b200_last_stage_ms(out))
    return tuple(out)


def last_plan():
    """Plan of the last moments call: dict(factorised, pair, diag, os, asym, k1_launches)."""
    out = (_i64 * 6)()
    check(_libagf2.agf2_b200_last_plan(out))
    return dict(zip(("factorised", "pair", "diag", "os", "asym", "k1_launches"), [int(x) for x in out]))


def plan_stats(nphys, nocc, nvir, naux, nocc_b=0, nvir_b=0, istart=0, iend=None, os_factor=1.0, ss_factor=1.0,
               uhf=False, part=0, nparts=1, cover=None):
    """Work plan of a moments call (host only, no GPU): dict of planner statistics; ``cover`` (int64,
    nocc x (nocc + nocc_b)) accumulates how many 8-column blocks of every (xi|ja) block this part forms."""
    iend = nocc if iend is None else iend
    out = (_i64 * 8)()
    cptr = None
    if cover is not None:
        assert cover.dtype == np.int64 and cover.flags.c_contiguous and cover.shape == (nocc, nocc + (nocc_b if uhf else 0))
        cptr = cover.ctypes.data_as(ctypes.POINTER(_i64))
    check(_libagf2.agf2_b200_plan_stats(nphys, nocc, nvir, naux, nocc_b, nvir_b, istart, iend, os_factor, ss_factor,
                                        1 if uhf else 0, part, nparts, cptr, out))
    keys = ("factorised", "chunks", "tiles", "thin_tiles", "max_cols", "cap_cols", "max_items", "blocks_per_pair")
    return dict(zip(keys, [int(x) for x in out]))


def set_coulomb_factorization(mode):
    """0 / False: off, 1 / True: always, 2: automatic (default)."""
    check(_libagf2.agf2_b200_set_coulomb_factorization(int(mode)))


def set_overlap(on):
    check(_libagf2.agf2_b200_set_overlap(1 if on else 0))


def set_workspace_bytes(nbytes):
    check(_libagf2.agf2_b200_set_workspace_bytes(int(nbytes)))


def set_device(index):
    """Make GPU ``index`` the calling thread's current device (the library's default context follows it)."""
    check(_libagf2.agf2_b200_set_device(int(index)))


def set_deterministic(on):
    """Stream-K partial tiles summed in a fixed order (default on): the same build gives the same bits."""
    check(_libagf2.agf2_b200_set_deterministic(1 if on else 0))


def set_option(name, value, ctx=None):
    check(_libagf2.agf2_b200_ctx_set_option(ctx, name.encode(), int(value)))


def sticky_status(clear=False):
    """(status, message) of the first failure of a ``void`` drop-in call (they cannot return one); 0 = none."""
    st = int(_libagf2.agf2_b200_sticky_status(1 if clear else 0))
    return st, (_libagf2.agf2_b200_last_error().decode() if st else "")


def flop_counters(reset=False, ctx=None):
    """FP64 tensor-core flops executed since the last reset: (k1_form, k2_moment, general GEMM), padding included."""
    out = (_dbl * 3)()
    check(_libagf2.agf2_b200_ctx_flop_counters(ctx, out, 1 if reset else 0))
    return tuple(out)


def last_comm_bytes(ctx=None):
    """Bytes this rank sent over NVLink in the last communicator-aware call: (all-gather, all-reduce)."""
    out = (_i64 * 2)()
    check(_libagf2.agf2_b200_ctx_last_comm_bytes(ctx, out))
    return int(out[0]), int(out[1])


# ---- communicator of the default context (one rank per process / GPU; NCCL inside the library) ----
def comm_unique_id():
    buf = ctypes.create_string_buffer(128)
    check(_libagf2.agf2_b200_comm_get_unique_id(buf))
    return buf.raw


def comm_init(uid, rank, size, ctx=None):
    assert len(uid) == 128
    check(_libagf2.agf2_b200_comm_init(ctx, ctypes.create_string_buffer(uid, 128), int(rank), int(size)))


def comm_info(ctx=None):
    r, n = ctypes.c_int(0), ctypes.c_int(1)
    check(_libagf2.agf2_b200_comm_info(ctx, ctypes.byref(r), ctypes.byref(n)))
    return r.value, n.value


def comm_barrier(ctx=None):
    check(_libagf2.agf2_b200_comm_barrier(ctx))


def comm_destroy(ctx=None):
    check(_libagf2.agf2_b200_comm_destroy(ctx))


def comm_allreduce(m, op="sum", ctx=None):
    """Sum (max / min) over the ranks of the communicator.  CUDA f64 tensors: in place, asynchronous on the current
    torch stream.  NumPy arrays: staged through the GPU, returns the reduced array."""
    import torch
    opc = {"sum": 0, "max": 1, "min": 2}[op]
    if isinstance(m, torch.Tensor):
        _check_dev(m, "tensor")
        check(_libagf2.agf2_b200_comm_allreduce(ctx, m.data_ptr(), m.numel(), opc,
                                                torch.cuda.current_stream(m.device).cuda_stream))
        return m
    a = np.ascontiguousarray(m, dtype=_f64)
    t = torch.from_numpy(a).cuda()
    check(_libagf2.agf2_b200_comm_allreduce(ctx, t.data_ptr(), t.numel(), opc, torch.cuda.current_stream().cuda_stream))
    return t.cpu().numpy().reshape(a.shape)


def _c(a, ndim):
    a = np.ascontiguousarray(a, dtype=_f64)
    assert a.ndim == ndim
    return a


def build_part_loop_rhf(ixq, qja, gf_occ, gf_vir, istart, iend, vv=None, vev=None,
                        os_factor=1.0, ss_factor=1.0, part=0, nparts=1):
    """Inner loop of ``OptRAGF2.build_part`` on the GPU (signature of auxgf/lib/agf2.py:63)."""
    nocc = gf_occ.naux
    nvir = gf_vir.naux
    nphys = gf_occ.nphys

    ei = _c(gf_occ.e, 1)
    ea = _c(gf_vir.e, 1)

    ixq = _c(np.asarray(ixq).reshape(nocc * nphys, -1), 2)
    qja = _c(np.asarray(qja).reshape(-1, nocc * nvir), 2)
    naux = qja.shape[0]

    if vv is None:
        vv = np.zeros((nphys, nphys), dtype=_f64, order="C")
    if vev is None:
        vev = np.zeros((nphys, nphys), dtype=_f64, order="C")
    vv = np.ascontiguousarray(vv)
    vev = np.ascontiguousarray(vev)

    check(_libagf2.agf2_b200_build_part_loop_rhf_sharded(ixq, qja, ei, ea, nphys, nocc, nvir, naux, istart, iend,
                                                         os_factor, ss_factor, part, nparts, vv, vev))
    return vv, vev


def build_part_loop_uhf(ixq, qja, gf_occ, gf_vir, istart, iend, vv=None, vev=None,
                        os_factor=1.0, ss_factor=1.0, part=0, nparts=1):
    """Inner loop of ``OptUAGF2.build_part`` for one spin channel (auxgf/lib/agf2.py:102)."""
    nocc_a = gf_occ[0].naux
    nocc_b = gf_occ[1].naux
    nvir_a = gf_vir[0].naux
    nvir_b = gf_vir[1].naux
    nphys = gf_occ[0].nphys

    ei_a = _c(gf_occ[0].e, 1)
    ei_b = _c(gf_occ[1].e, 1)
    ea_a = _c(gf_vir[0].e, 1)
    ea_b = _c(gf_vir[1].e, 1)

    ixq_a = _c(np.asarray(ixq[0]).reshape(nocc_a * nphys, -1), 2)
    qja_a = _c(np.asarray(qja[0]).reshape(-1, nocc_a * nvir_a), 2)
    naux = qja_a.shape[0]
    qja_b = _c(np.asarray(qja[1]).reshape(naux, nocc_b * nvir_b), 2)

    if vv is None:
        vv = np.zeros((nphys, nphys), dtype=_f64, order="C")
    if vev is None:
        vev = np.zeros((nphys, nphys), dtype=_f64, order="C")
    vv = np.ascontiguousarray(vv)
    vev = np.ascontiguousarray(vev)

    check(_libagf2.agf2_b200_build_part_loop_uhf_sharded(ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b, nphys, nocc_a,
                                                         nocc_b, nvir_a, nvir_b, naux, istart, iend, os_factor, ss_factor,
                                                         part, nparts, vv, vev))
    return vv, vev


# ---- raw-array conveniences (no Aux objects) --------------------------------------------------
class _Poles:
    def __init__(self, e, nphys):
        self.e = np.asarray(e, dtype=_f64)
        self.naux = self.e.size
        self.nphys = nphys


def moments_rhf(ixq, qja, ei, ea, nphys, istart=0, iend=None, os_factor=1.0, ss_factor=1.0, part=0, nparts=1):
    occ, vir = _Poles(ei, nphys), _Poles(ea, nphys)
    return build_part_loop_rhf(ixq, qja, occ, vir, istart, occ.naux if iend is None else iend,
                               os_factor=os_factor, ss_factor=ss_factor, part=part, nparts=nparts)


def moments_uhf(ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b, nphys, istart=0, iend=None,
                os_factor=1.0, ss_factor=1.0):
    occ = (_Poles(ei_a, nphys), _Poles(ei_b, nphys))
    vir = (_Poles(ea_a, nphys), _Poles(ea_b, nphys))
    return build_part_loop_uhf((ixq_a, None), (qja_a, qja_b), occ, vir, istart,
                               occ[0].naux if iend is None else iend, os_factor=os_factor, ss_factor=ss_factor)


def compress(vv, vev):
    """Device pole compression (auxgf/agf2/optragf2.py:270-278) on host arrays -> (e, c)."""
    vv = _c(vv, 2); vev = _c(vev, 2)
    n = vv.shape[0]
    e = np.empty(n); c = np.empty((n, n))
    check(_libagf2.agf2_b200_compress(vv.ctypes.data, vev.ctypes.data, n, e.ctypes.data, c.ctypes.data, 0))
    return e, c


def eigh(a):
    """cuSOLVER dsyevd on a host symmetric matrix (auxgf/aux/eig.py:23-33) -> (w, v)."""
    a = _c(a, 2)
    n = a.shape[0]
    w = np.empty(n); v = np.empty((n, n))
    check(_libagf2.agf2_b200_eigh(a.ctypes.data, n, w.ctypes.data, v.ctypes.data, 0))
    return w, v


# ---- device-resident calls (torch tensors as containers only) ---------------------------------
def _check_dev(t, name):
    import torch
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise ValueError("%s must be a contiguous float64 CUDA tensor" % name)
    return t


def moments_rhf_dev(ixq, qja, ei, ea, istart=0, iend=None, os_factor=1.0, ss_factor=1.0, part=0, nparts=1,
                    vv=None, vev=None, sync=True, naux=None, comm=False, ei_host=None):
    """Device-resident moment build.  ixq: (nocc, nphys, ld>=naux), qja: (naux, nocc, ld>=nvir) CUDA f64
    tensors whose last dimension may be padded to an even pitch (logical sizes: ``naux`` keyword,
    ``ea.numel()``).  Returns (vv, vev) CUDA tensors (accumulated into when given).
    ``comm=True``: this rank's share of the pair-work items (rank / size of the library's communicator) followed by
    ONE NCCL all-reduce of the packed [vv|vev] inside the library, on the same stream.  ``ei_host``: NumPy copy of
    ``ei`` (saves the library a device->host copy and a stream synchronisation)."""
    import torch
    _check_dev(ixq, "ixq"); _check_dev(qja, "qja"); _check_dev(ei, "ei"); _check_dev(ea, "ea")
    nocc, nphys, ld_ixq = ixq.shape
    naux = ld_ixq if naux is None else naux
    ld_qja = qja.shape[2]
    nvir = ea.numel()
    assert qja.shape[0] == naux and qja.shape[1] == nocc and ei.numel() == nocc and ld_qja >= nvir
    iend = nocc if iend is None else iend
    if vv is None:
        vv = torch.zeros((nphys, nphys), dtype=torch.float64, device=ixq.device)
    if vev is None:
        vev = torch.zeros((nphys, nphys), dtype=torch.float64, device=ixq.device)
    stream = torch.cuda.current_stream(ixq.device).cuda_stream
    if comm:
        eh = None if ei_host is None else _c(ei_host, 1)
        check(_libagf2.agf2_b200_ctx_moments_rhf_dev(None, ixq.data_ptr(), ld_ixq, qja.data_ptr(), ld_qja,
                                                     ei.data_ptr(), ea.data_ptr(), None if eh is None else eh.ctypes.data,
                                                     nphys, nocc, nvir, naux, istart, iend, os_factor, ss_factor,
                                                     vv.data_ptr(), vev.data_ptr(), stream, 1 if sync else 0))
    else:
        check(_libagf2.agf2_b200_moments_rhf_dev(ixq.data_ptr(), ld_ixq, qja.data_ptr(), ld_qja, ei.data_ptr(),
                                                 ea.data_ptr(), nphys, nocc, nvir, naux, istart, iend, os_factor,
                                                 ss_factor, part, nparts, vv.data_ptr(), vev.data_ptr(), stream,
                                                 1 if sync else 0))
    return vv, vev


def pole_block(nocc, rank, size):
    """[lo, hi): the poles of (ix|Q) that rank ``rank`` of ``size`` holds in a pole-sharded build."""
    lo, hi = _i64(0), _i64(0)
    check(_libagf2.agf2_b200_pole_block(nocc, rank, size, ctypes.byref(lo), ctypes.byref(hi)))
    return int(lo.value), int(hi.value)


def shard_sub_block(nocc, nphys, ld_ixq, nranks, ctx=None):
    """Poles per visiting sub-block of a pole-sharded build of this shape (``visiting_bytes`` option)."""
    return int(_libagf2.agf2_b200_ctx_shard_sub_block(ctx, nocc, nphys, ld_ixq, nranks))


def shard_cover(nocc, nranks, sub_block, limit=0):
    """Host-only replay of the pole-sharded pair schedule: (cover, stats) with cover[i, j] = how often the ranks form
    the unordered pair {i, j} (i > j; must be exactly 1) and stats = per-rank (pairs, poles sent), shape (nranks, 2)."""
    cover = np.zeros((nocc, nocc), dtype=np.int64)
    stats = np.zeros((nranks, 2), dtype=np.int64)
    check(_libagf2.agf2_b200_shard_cover(nocc, nranks, sub_block, limit, cover.ctypes.data_as(ctypes.POINTER(_i64)),
                                         stats.ctypes.data_as(ctypes.POINTER(_i64))))
    return cover, stats


def last_p2p_bytes(ctx=None):
    """Bytes this rank sent point-to-point over NVLink (visiting pole blocks) in the last pole-sharded build."""
    return int(_libagf2.agf2_b200_ctx_last_p2p_bytes(ctx))


def moments_rhf_dev_sharded(ixq_block, qja, ei, ea, os_factor=1.0, ss_factor=1.0, vv=None, vev=None, sync=True,
                            naux=None, ei_host=None, ctx=None):
    """Pole-sharded device-resident build on the library's communicator: ``ixq_block`` holds only this rank's poles
    ``pole_block(nocc, rank, size)`` of (ix|Q) -- (hi - lo, nphys, ld >= naux) -- while ``qja``, ``ei``, ``ea`` are
    complete on every rank; the other ranks' poles visit over NVLink inside the call and one all-reduce gives every
    rank the full (vv, vev).  For shapes whose (ix|Q) does not fit one GPU (BASELINE configs[4])."""
    import torch
    _check_dev(ixq_block, "ixq_block"); _check_dev(qja, "qja"); _check_dev(ei, "ei"); _check_dev(ea, "ea")
    nloc, nphys, ld_ixq = ixq_block.shape
    naux = ld_ixq if naux is None else naux
    nocc, nvir, ld_qja = ei.numel(), ea.numel(), qja.shape[2]
    rank, size = comm_info(ctx)
    lo, hi = pole_block(nocc, rank, size)
    assert nloc == hi - lo, "ixq_block must hold the poles [%d, %d) of this rank" % (lo, hi)
    assert qja.shape[0] == naux and qja.shape[1] == nocc and ld_qja >= nvir
    if vv is None:
        vv = torch.zeros((nphys, nphys), dtype=torch.float64, device=qja.device)
    if vev is None:
        vev = torch.zeros((nphys, nphys), dtype=torch.float64, device=qja.device)
    eh = None if ei_host is None else _c(ei_host, 1)
    check(_libagf2.agf2_b200_ctx_moments_rhf_dev_sharded(
        ctx, ixq_block.data_ptr(), ld_ixq, qja.data_ptr(), ld_qja, ei.data_ptr(), ea.data_ptr(),
        None if eh is None else eh.ctypes.data, nphys, nocc, nvir, naux, os_factor, ss_factor, vv.data_ptr(),
        vev.data_ptr(), torch.cuda.current_stream(qja.device).cuda_stream, 1 if sync else 0))
    return vv, vev


def moments_uhf_dev(ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b, istart=0, iend=None, os_factor=1.0,
                    ss_factor=1.0, part=0, nparts=1, vv=None, vev=None, sync=True, naux=None, comm=False):
    import torch
    for t, n in ((ixq_a, "ixq_a"), (qja_a, "qja_a"), (qja_b, "qja_b"), (ei_a, "ei_a"), (ei_b, "ei_b"),
                 (ea_a, "ea_a"), (ea_b, "ea_b")):
        _check_dev(t, n)
    nocc_a, nphys, ld_ixq = ixq_a.shape
    naux = ld_ixq if naux is None else naux
    nocc_b, nvir_a, nvir_b = ei_b.numel(), ea_a.numel(), ea_b.numel()
    iend = nocc_a if iend is None else iend
    if vv is None:
        vv = torch.zeros((nphys, nphys), dtype=torch.float64, device=ixq_a.device)
    if vev is None:
        vev = torch.zeros((nphys, nphys), dtype=torch.float64, device=ixq_a.device)
    stream = torch.cuda.current_stream(ixq_a.device).cuda_stream
    if comm:
        check(_libagf2.agf2_b200_ctx_moments_uhf_dev(None, ixq_a.data_ptr(), ld_ixq, qja_a.data_ptr(), qja_a.shape[2],
                                                     qja_b.data_ptr(), qja_b.shape[2], ei_a.data_ptr(), ei_b.data_ptr(),
                                                     ea_a.data_ptr(), ea_b.data_ptr(), None, None, nphys, nocc_a, nocc_b,
                                                     nvir_a, nvir_b, naux, istart, iend, os_factor, ss_factor,
                                                     vv.data_ptr(), vev.data_ptr(), stream, 1 if sync else 0))
    else:
        check(_libagf2.agf2_b200_moments_uhf_dev(ixq_a.data_ptr(), ld_ixq, qja_a.data_ptr(), qja_a.shape[2],
                                                 qja_b.data_ptr(), qja_b.shape[2], ei_a.data_ptr(), ei_b.data_ptr(),
                                                 ea_a.data_ptr(), ea_b.data_ptr(), nphys, nocc_a, nocc_b, nvir_a,
                                                 nvir_b, naux, istart, iend, os_factor, ss_factor, part, nparts,
                                                 vv.data_ptr(), vev.data_ptr(), stream, 1 if sync else 0))
    return vv, vev


def moments_rhf_multi(ixq, qja, ei, ea, nphys, os_factor=1.0, ss_factor=1.0):
    """Host arrays in, host (vv, vev) out, on every rank of the library's communicator (one process per GPU) -- the
    multi-GPU counterpart of ``moments_rhf`` and the call that replaces the replicated MPI ranks + Reduce/Bcast of
    auxgf/mpi/agf2/optragf2.py:370-395.  Everything happens inside ``agf2_b200_ctx_build_part_loop_rhf``: rank r copies
    1/N of ``ixq`` and ``qja`` over PCIe, in groups pipelined under the pair loop; the GPUs all-gather the groups over
    NVLink; each builds its share of the pair-work items; ONE all-reduce sums the packed [vv|vev]."""
    ei = _c(ei, 1); ea = _c(ea, 1)
    nocc, nvir = ei.size, ea.size
    ixq = _c(np.asarray(ixq).reshape(nocc * nphys, -1), 2)
    naux = ixq.shape[1]
    qja = _c(np.asarray(qja).reshape(naux, nocc * nvir), 2)
    vv = np.zeros((nphys, nphys)); vev = np.zeros((nphys, nphys))
    check(_libagf2.agf2_b200_ctx_build_part_loop_rhf(None, ixq, qja, ei, ea, nphys, nocc, nvir, naux, 0, nocc,
                                                     os_factor, ss_factor, vv, vev))
    return vv, vev


def moments_uhf_multi(ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b, nphys, os_factor=1.0, ss_factor=1.0):
    """UHF channel counterpart of ``moments_rhf_multi`` (``agf2_b200_ctx_build_part_loop_uhf``)."""
    ei_a = _c(ei_a, 1); ei_b = _c(ei_b, 1); ea_a = _c(ea_a, 1); ea_b = _c(ea_b, 1)
    oa, ob, va, vb = ei_a.size, ei_b.size, ea_a.size, ea_b.size
    ixq_a = _c(np.asarray(ixq_a).reshape(oa * nphys, -1), 2)
    naux = ixq_a.shape[1]
    qja_a = _c(np.asarray(qja_a).reshape(naux, oa * va), 2)
    qja_b = _c(np.asarray(qja_b).reshape(naux, ob * vb), 2)
    vv = np.zeros((nphys, nphys)); vev = np.zeros((nphys, nphys))
    check(_libagf2.agf2_b200_ctx_build_part_loop_uhf(None, ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b, nphys, oa, ob,
                                                     va, vb, naux, 0, oa, os_factor, ss_factor, vv, vev))
    return vv, vev


def gemm_dev(a, b, out=None, transpose=False, accumulate=False, ctx=None):
    """out = a @ b.T (``transpose``: out holds the transpose) on the library's DMMA pipeline.  a: (m, k), b: (n, k)
    CUDA f64 tensors, last dimension contiguous, even row pitch."""
    import torch
    m, k = a.shape
    n = b.shape[0]
    assert b.shape[1] == k and a.stride(1) == 1 and b.stride(1) == 1
    if out is None:
        out = torch.empty((n, m) if transpose else (m, n), dtype=torch.float64, device=a.device)
    check(_libagf2.agf2_b200_ctx_gemm(ctx, a.data_ptr(), m, a.stride(0), b.data_ptr(), n, b.stride(0), k,
                                      out.data_ptr(), out.stride(0), 1 if transpose else 0, 1 if accumulate else 0,
                                      torch.cuda.current_stream(a.device).cuda_stream))
    return out


def host_register(a):
    """Page-lock a NumPy array so that the host-pointer calls upload it by asynchronous DMA (optional)."""
    check(_libagf2.agf2_b200_host_register(a.ctypes.data, a.nbytes))


def host_unregister(a):
    check(_libagf2.agf2_b200_host_unregister(a.ctypes.data))


_libagf2.agf2_b200_fock_create.argtypes = [ctypes.POINTER(_vp)]
_libagf2.agf2_b200_fock_destroy.argtypes = [_vp]
_libagf2.agf2_b200_fock_set_couplings.argtypes = [_vp, _vp, _i64, _i64]
_libagf2.agf2_b200_fock_solve.argtypes = [_vp, _vp, _vp, _dbl, _vp, _vp]
_libagf2.agf2_b200_fock_solve_async.argtypes = [_vp, _vp, _vp, _dbl, _vp, _vp]
_libagf2.agf2_b200_fock_wait.argtypes = [_vp]
_libagf2.agf2_b200_energy_2body.argtypes = [_vp, _vp, _i64, _vp, _vp, _i64, _i64, ctypes.POINTER(_dbl)]
SYMBOLS += ["agf2_b200_fock_create", "agf2_b200_fock_destroy", "agf2_b200_fock_set_couplings", "agf2_b200_fock_solve",
            "agf2_b200_fock_solve_async", "agf2_b200_fock_wait", "agf2_b200_energy_2body"]


class FockSolver:
    """Eigensolver of the extended Fock matrix [[F, v], [v^T, diag(e - shift)]] with the couplings ``v`` resident in
    HBM (``agf2_b200_fock_*``): the matrix is assembled on the device, cuSOLVER dsyevd runs in a reused workspace and
    only ``w`` and the physical rows of the eigenvectors come back.  Replaces np.linalg.eigh of auxgf/aux/eig.py:23-33
    inside the Fock loop (auxgf/agf2/fock.py:119-143, chempot.py:11-49)."""

    def __init__(self):
        h = _vp()
        check(_libagf2.agf2_b200_fock_create(ctypes.byref(h)))
        self._h = h
        self._v = None
        self._out = None

    def set_couplings(self, v):
        v = _c(v, 2)
        self._v = v                     # keeps the identity the cache in Aux.eig_phys compares
        self.nphys, self.naux = v.shape
        check(_libagf2.agf2_b200_fock_set_couplings(self._h, v.ctypes.data, self.nphys, self.naux))

    def solve_async(self, fock, e, shift=0.0):
        fock = _c(fock, 2); e = _c(e, 1)
        assert fock.shape == (self.nphys, self.nphys) and e.size == self.naux
        n = self.nphys + self.naux
        w = np.empty(n); v = np.empty((self.nphys, n))
        check(_libagf2.agf2_b200_fock_solve_async(self._h, fock.ctypes.data, e.ctypes.data, float(shift), w.ctypes.data,
                                                  v.ctypes.data))
        self._out = (w, v)

    def wait(self):
        check(_libagf2.agf2_b200_fock_wait(self._h))
        out, self._out = self._out, None
        return out

    def solve(self, fock, e, shift=0.0):
        self.solve_async(fock, e, shift)
        return self.wait()

    def close(self):
        if getattr(self, "_h", None):
            _libagf2.agf2_b200_fock_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def energy_2body(gv, ge, sv, se):
    """sum_{l,k} (sum_x gv[x,l] sv[x,k])^2 / (ge[l] - se[k]) on the device; gv (nphys, nl), sv (nphys, nk) as in
    the reference's Aux couplings (auxgf/aux/energy.py:43-49)."""
    gvt = _c(np.asarray(gv).T, 2); svt = _c(np.asarray(sv).T, 2)
    ge = _c(ge, 1); se = _c(se, 1)
    out = _dbl(0.0)
    check(_libagf2.agf2_b200_energy_2body(gvt.ctypes.data, ge.ctypes.data, ge.size, svt.ctypes.data, se.ctypes.data,
                                          se.size, gvt.shape[1] if ge.size else svt.shape[1], ctypes.byref(out)))
    return out.value


def compress_dev(vv, vev):
    import torch
    n = vv.shape[0]
    e = torch.empty(n, dtype=torch.float64, device=vv.device)
    c = torch.empty((n, n), dtype=torch.float64, device=vv.device)
    torch.cuda.current_stream(vv.device).synchronize()
    check(_libagf2.agf2_b200_compress(vv.data_ptr(), vev.data_ptr(), n, e.data_ptr(), c.data_ptr(), 1))
    return e, c


def eigh_dev(a):
    import torch
    n = a.shape[0]
    w = torch.empty(n, dtype=torch.float64, device=a.device)
    v = torch.empty((n, n), dtype=torch.float64, device=a.device)
    torch.cuda.current_stream(a.device).synchronize()
    check(_libagf2.agf2_b200_eigh(a.data_ptr(), n, w.data_ptr(), v.data_ptr(), 1))
    return w, v


# ---- device-resident DF tensor + fused sector build ---------------------------------------------
_libagf2.agf2_b200_ao2mo_last_error.restype = ctypes.c_char_p
_libagf2.agf2_b200_eri_create.argtypes = [_vp, _i64, _i64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp)]
_libagf2.agf2_b200_eri_destroy.argtypes = [_vp]
_libagf2.agf2_b200_eri_full_ptr.argtypes = [_vp, ctypes.POINTER(_vp)]
_libagf2.agf2_b200_ao2mo_sector.argtypes = [_vp, _vp, _i64, _vp, _i64, ctypes.c_int, _vp, ctypes.POINTER(_vp),
                                            ctypes.POINTER(_i64), ctypes.POINTER(_vp), ctypes.POINTER(_i64)]
_libagf2.agf2_b200_get_jk.argtypes = [_vp, _vp, _vp, _vp, ctypes.c_int, _vp]
SYMBOLS += ["agf2_b200_eri_create", "agf2_b200_eri_destroy", "agf2_b200_eri_full_ptr", "agf2_b200_ao2mo_sector",
            "agf2_b200_get_jk"]


def _check2(status):
    if status != 0:
        msg = _libagf2.agf2_b200_ao2mo_last_error().decode() or _libagf2.agf2_b200_last_error().decode()
        raise Agf2B200Error("libagf2_b200 error %d: %s" % (status, msg))


class DeviceERI:
    """The DF tensor L (naux, nphys, nphys) staged once to HBM, plus the per-sector pipeline
    ao2mo -> moment build on the device (replaces optragf2.py:243-268 with no host round trip of
    ixq/qja).  ``eri`` is a host NumPy array: tril-packed (naux, nphys(nphys+1)/2) or (naux, nphys, nphys)."""

    def __init__(self, eri, nphys=None, packed=None):
        eri = np.ascontiguousarray(eri, dtype=_f64)
        if packed is None:
            packed = eri.ndim == 2
        if packed:
            naux = eri.shape[0]
            if nphys is None:
                nphys = int((np.sqrt(8 * eri.shape[1] + 1) - 1) / 2)
            assert eri.shape[1] == nphys * (nphys + 1) // 2
        else:
            eri = eri.reshape(eri.shape[0], -1)
            naux = eri.shape[0]
            nphys = int(round(np.sqrt(eri.shape[1]))) if nphys is None else nphys
            assert eri.shape[1] == nphys * nphys
        self.naux, self.nphys = naux, nphys
        h = _vp()
        _check2(_libagf2.agf2_b200_eri_create(eri.ctypes.data, naux, nphys, 1 if packed else 0, 0, ctypes.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            _libagf2.agf2_b200_eri_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        import torch
        return torch.cuda.current_stream().cuda_stream

    def sector(self, v_occ, v_vir):
        """(ix|Q), (Q|ja) for one sector -> raw device pointers (ixq, ld_ixq, qja, ld_qja)."""
        v_occ = _c(v_occ, 2); v_vir = _c(v_vir, 2)
        assert v_occ.shape[0] == self.nphys and v_vir.shape[0] == self.nphys
        ixq, qja, ldi, ldq = _vp(), _vp(), _i64(), _i64()
        _check2(_libagf2.agf2_b200_ao2mo_sector(self._h, v_occ.ctypes.data, v_occ.shape[1], v_vir.ctypes.data,
                                                v_vir.shape[1], 0, self._stream(), ctypes.byref(ixq),
                                                ctypes.byref(ldi), ctypes.byref(qja), ctypes.byref(ldq)))
        return ixq.value, ldi.value, qja.value, ldq.value

    def get_jk(self, rdm1):
        """J, K (host arrays) of a host density matrix, built on the device from the resident DF tensor."""
        rdm1 = _c(rdm1, 2)
        j = np.empty_like(rdm1); k = np.empty_like(rdm1)
        _check2(_libagf2.agf2_b200_get_jk(self._h, rdm1.ctypes.data, j.ctypes.data, k.ctypes.data, 0, self._stream()))
        return j, k

    def moments_rhf(self, gf_occ, gf_vir, os_factor=1.0, ss_factor=1.0, part=None, nparts=None):
        """Packed [vv|vev] (CUDA tensor) of one sector, everything on the device.  By default the work is shared over
        the ranks of the library's communicator and summed with one all-reduce inside the library; explicit
        ``part``/``nparts`` return this part's contribution only."""
        import torch
        use_comm = part is None
        part, nparts = (0, 1) if part is None else (part, nparts)
        dev = torch.device("cuda", torch.cuda.current_device())
        nphys, nocc, nvir = self.nphys, gf_occ.naux, gf_vir.naux
        ixq, ldi, qja, ldq = self.sector(gf_occ.v, gf_vir.v)
        ei = torch.from_numpy(_c(gf_occ.e, 1)).to(dev)
        ea = torch.from_numpy(_c(gf_vir.e, 1)).to(dev)
        out = torch.zeros((2, nphys, nphys), dtype=torch.float64, device=dev)
        if use_comm:
            check(_libagf2.agf2_b200_ctx_moments_rhf_dev(None, ixq, ldi, qja, ldq, ei.data_ptr(), ea.data_ptr(),
                                                         _c(gf_occ.e, 1).ctypes.data, nphys, nocc, nvir, self.naux, 0,
                                                         nocc, os_factor, ss_factor, out[0].data_ptr(),
                                                         out[1].data_ptr(), self._stream(), 1))
        else:
            check(_libagf2.agf2_b200_moments_rhf_dev(ixq, ldi, qja, ldq, ei.data_ptr(), ea.data_ptr(), nphys, nocc,
                                                     nvir, self.naux, 0, nocc, os_factor, ss_factor, part, nparts,
                                                     out[0].data_ptr(), out[1].data_ptr(), self._stream(), 1))
        return out

    @staticmethod
    def compress(mom):
        """cuSOLVER pole compression of a packed [vv|vev] CUDA tensor -> host (e, c)."""
        e, c = compress_dev(mom[0], mom[1])
        return e.cpu().numpy(), c.cpu().numpy()

    @staticmethod
    def moments_uhf(eri_a, eri_b, gf_occ, gf_vir, os_factor=1.0, ss_factor=1.0, part=None, nparts=None):
        """One spin channel of OptUAGF2.build_part: eri_a/gf_*[0] same spin, eri_b/gf_*[1] opposite spin.
        ``part``/``nparts`` as in ``moments_rhf``."""
        import torch
        if eri_a is eri_b:
            raise ValueError("moments_uhf needs one staged DeviceERI per spin: sector() scratch of a handle is "
                             "overwritten by its next sector() call")
        use_comm = part is None
        part, nparts = (0, 1) if part is None else (part, nparts)
        dev = torch.device("cuda", torch.cuda.current_device())
        nphys = eri_a.nphys
        ixq, ldi, qja, ldq = eri_a.sector(gf_occ[0].v, gf_vir[0].v)
        _, _, qjb, ldqb = eri_b.sector(gf_occ[1].v, gf_vir[1].v)
        T = lambda a: torch.from_numpy(_c(a, 1)).to(dev)  # noqa: E731
        eia, eib, eaa, eab = T(gf_occ[0].e), T(gf_occ[1].e), T(gf_vir[0].e), T(gf_vir[1].e)
        out = torch.zeros((2, nphys, nphys), dtype=torch.float64, device=dev)
        if use_comm:
            check(_libagf2.agf2_b200_ctx_moments_uhf_dev(None, ixq, ldi, qja, ldq, qjb, ldqb, eia.data_ptr(),
                                                         eib.data_ptr(), eaa.data_ptr(), eab.data_ptr(),
                                                         _c(gf_occ[0].e, 1).ctypes.data, _c(gf_occ[1].e, 1).ctypes.data,
                                                         nphys, gf_occ[0].naux, gf_occ[1].naux, gf_vir[0].naux,
                                                         gf_vir[1].naux, eri_a.naux, 0, gf_occ[0].naux, os_factor,
                                                         ss_factor, out[0].data_ptr(), out[1].data_ptr(),
                                                         eri_a._stream(), 1))
        else:
            check(_libagf2.agf2_b200_moments_uhf_dev(ixq, ldi, qja, ldq, qjb, ldqb, eia.data_ptr(), eib.data_ptr(),
                                                     eaa.data_ptr(), eab.data_ptr(), nphys, gf_occ[0].naux,
                                                     gf_occ[1].naux, gf_vir[0].naux, gf_vir[1].naux, eri_a.naux, 0,
                                                     gf_occ[0].naux, os_factor, ss_factor, part, nparts,
                                                     out[0].data_ptr(), out[1].data_ptr(), eri_a._stream(), 1))
        return out
