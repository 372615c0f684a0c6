# -*- coding: utf-8 -*-
"""
ctypes binding of ``libaucell_b200.so`` (C-ABI declared in ``include/aucell_b200.h``) and its in-tree build.

There is no CPU fallback: if the library cannot be loaded, or no sm_100 device is present, every compute entry
point raises.  The shared object is built in-tree (``pyscenic_b200/libaucell_b200.so``) with
``nvcc -gencode arch=compute_100a,code=sm_100a``.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess
import threading
from typing import List, Optional

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG_DIR)
_CSRC = os.path.join(_PKG_DIR, "csrc")
_INCLUDE = os.path.join(_ROOT, "include")
LIB_NAME = "libaucell_b200.so"
LIB_PATH = os.path.join(_PKG_DIR, LIB_NAME)

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
]
# translation units of the library (compiled in parallel, then linked into one shared object)
TRANSLATION_UNITS = ["aucell_b200.cu", "aucell_f32_dense.cu", "aucell_f32_csr.cu", "aucell_f64_dense.cu",
                     "aucell_f64_csr.cu", "aucell_f32_csr16.cu", "aucell_rss.cu", "aucell_binarize.cu"]

# every symbol include/aucell_b200.h declares (tests check that the library exports each of them)
EXPORTED_SYMBOLS = [
    "aucell_b200_abi_version",
    "aucell_b200_last_error",
    "aucell_b200_device_info",
    "aucell_b200_max_genes",
    "aucell_plan_create",
    "aucell_plan_destroy",
    "aucell_rank_dense_f32",
    "aucell_rank_dense_f64",
    "aucell_rank_csr_f32",
    "aucell_rank_csr_f64",
    "aucell_dense_f32",
    "aucell_dense_f64",
    "aucell_csr_f32",
    "aucell_csr_f64",
    "aucell_csr16_f32",
    "aucell_from_rankings_u32",
    "aucell_csr_f32_host",
    "aucell_dense_f32_host",
    "aucell_csr_is_canonical",
    "aucell_b200_release_staging",
    "aucell_dense_strided_host",
    "aucell_cast_f64_to_f32",
    "aucell_cast_i64_to_i32",
    "aucell_host_dense_to_csr",
    "aucell_b200_launch_count",
    "aucell_plan_path_stats",
    "aucell_plan_last_host_bytes",
    "aucell_rss_f64",
    "aucell_count_nonzero_csr_f32",
    "aucell_binarize_stats_f64",
    "aucell_binarize_dip_f64",
    "aucell_binarize_gmm2_f64",
    "aucell_binarize_trough_f64",
    "aucell_csr16_lut8_f32",
    "aucell_csr_lut8_f32",
    "aucell_plan_set_fast_path",
    "aucell_plan_fast_stats",
]


class AUCellNativeError(RuntimeError):
    """A call into libaucell_b200.so failed (the message is the library's own)."""

    def __init__(self, code: int, message: str):
        super().__init__("libaucell_b200: error {}: {}".format(code, message))
        self.code = code


def _sources() -> List[str]:
    return sorted(
        os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh", ".h"))
    ) + [os.path.join(_INCLUDE, "aucell_b200.h")]


def _nvcc() -> Optional[str]:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


_HASH_PATH = LIB_PATH + ".src.sha256"
_LOCK_PATH = LIB_PATH + ".lock"


def _source_hash() -> str:
    """sha256 over the sources and the compiler flags the library is built from."""
    import hashlib

    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS + TRANSLATION_UNITS).encode())
    for path in _sources():
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def is_stale() -> bool:
    """The library is rebuilt when the sources it was built from changed.  The decision is made on CONTENT (a hash
    written next to the library by build()), not on modification times: the library travels to other machines by
    file copy, which does not keep mtimes in order, and N ranks of one job must not all decide to rebuild."""
    if not os.path.exists(LIB_PATH):
        return True
    if os.path.exists(_HASH_PATH):
        with open(_HASH_PATH) as fh:
            return fh.read().strip() != _source_hash()
    t = os.path.getmtime(LIB_PATH)          # library built by hand (no hash file): fall back to mtimes
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a in-tree.  Cross-compiles without a GPU.  The translation units are
    compiled concurrently (object files under pyscenic_b200/build/) and linked with ``nvcc -shared``.  Safe to call
    from several processes at once (ranks of one job): a file lock serialises them and the late-comers find the
    library up to date."""
    import fcntl

    if not force and not is_stale():
        return LIB_PATH
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build {} (set NVCC or add nvcc to PATH)".format(LIB_NAME))
    with open(_LOCK_PATH, "w") as lock_fh:
        fcntl.flock(lock_fh, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():      # another process built it while we waited
                return LIB_PATH
            return _build_locked(nvcc, verbose)
        finally:
            fcntl.flock(lock_fh, fcntl.LOCK_UN)


def _build_locked(nvcc: str, verbose: bool) -> str:
    src_hash = _source_hash()
    obj_dir = os.path.join(_PKG_DIR, "build")
    os.makedirs(obj_dir, exist_ok=True)
    procs = []
    for tu in TRANSLATION_UNITS:
        obj = os.path.join(obj_dir, tu.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas=-v"] if verbose else []) + [
            "-I", _INCLUDE, "-I", _CSRC, "-c", "-o", obj, os.path.join(_CSRC, tu)]
        procs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for cmd, obj, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed:\n{}\n{}".format(" ".join(cmd), out))
        if verbose:
            print(out)
        objs.append(obj)
    tmp = "{}.tmp.{}".format(LIB_PATH, os.getpid())
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n{}\n{}".format(" ".join(link), res.stdout))
    os.replace(tmp, LIB_PATH)
    with open(_HASH_PATH + ".tmp", "w") as fh:
        fh.write(src_hash + "\n")
    os.replace(_HASH_PATH + ".tmp", _HASH_PATH)
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def load(auto_build: bool = True) -> ctypes.CDLL:
    """Load the library (building it first when sources are newer and nvcc is available)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if auto_build and is_stale() and _nvcc() is not None:
            build()
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "{} is missing and could not be built; pyscenic_b200 has no CPU fallback. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'` on a machine with nvcc.".format(LIB_PATH)
            )
        lib = ctypes.CDLL(LIB_PATH)
        c_int, i32, i64 = ctypes.c_int, ctypes.c_int32, ctypes.c_int64
        p = ctypes.c_void_p
        lib.aucell_b200_abi_version.argtypes = []
        lib.aucell_b200_abi_version.restype = c_int
        lib.aucell_b200_last_error.argtypes = []
        lib.aucell_b200_last_error.restype = ctypes.c_char_p
        lib.aucell_b200_device_info.argtypes = [c_int, p, p, p, p]
        lib.aucell_b200_device_info.restype = c_int
        lib.aucell_b200_max_genes.argtypes = []
        lib.aucell_b200_max_genes.restype = i64
        lib.aucell_b200_launch_count.argtypes = []
        lib.aucell_b200_launch_count.restype = i64
        lib.aucell_plan_create.argtypes = [c_int, i64, p, i32, p, p, p, p, p, i64, p]
        lib.aucell_plan_create.restype = c_int
        lib.aucell_plan_path_stats.argtypes = [p, p, c_int]
        lib.aucell_plan_path_stats.restype = c_int
        lib.aucell_plan_last_host_bytes.argtypes = [p, p]
        lib.aucell_plan_last_host_bytes.restype = c_int
        lib.aucell_plan_set_fast_path.argtypes = [p, c_int, p]
        lib.aucell_plan_set_fast_path.restype = c_int
        lib.aucell_plan_fast_stats.argtypes = [p, p, c_int]
        lib.aucell_plan_fast_stats.restype = c_int
        lib.aucell_rss_f64.argtypes = [c_int, p, i64, i64, i32, p, i32, p, p, p]
        lib.aucell_rss_f64.restype = c_int
        lib.aucell_count_nonzero_csr_f32.argtypes = [c_int, p, p, i64, p, p]
        lib.aucell_count_nonzero_csr_f32.restype = c_int
        f64 = ctypes.c_double
        lib.aucell_binarize_stats_f64.argtypes = [c_int, p, i32, i64, p, p]
        lib.aucell_binarize_dip_f64.argtypes = [c_int, p, i32, i64, p, p, p]
        lib.aucell_binarize_gmm2_f64.argtypes = [c_int, p, i32, i64, p, i32, f64, f64, p, p]
        lib.aucell_binarize_trough_f64.argtypes = [c_int, p, i32, i64, p, p, p, f64, i32, p, p]
        for name in ("aucell_binarize_stats_f64", "aucell_binarize_dip_f64", "aucell_binarize_gmm2_f64",
                     "aucell_binarize_trough_f64"):
            getattr(lib, name).restype = c_int
        lib.aucell_plan_destroy.argtypes = [p]
        lib.aucell_plan_destroy.restype = c_int
        for name in ("aucell_rank_dense_f32", "aucell_rank_dense_f64"):
            f = getattr(lib, name)
            f.argtypes = [p, p, i64, i64, p, p, p]
            f.restype = c_int
        for name in ("aucell_rank_csr_f32", "aucell_rank_csr_f64"):
            f = getattr(lib, name)
            f.argtypes = [p, p, p, p, i64, p, p, p]
            f.restype = c_int
        for name in ("aucell_dense_f32", "aucell_dense_f64"):
            f = getattr(lib, name)
            f.argtypes = [p, p, i64, i64, p, p, p]
            f.restype = c_int
        for name in ("aucell_csr_f32", "aucell_csr_f64", "aucell_csr16_f32"):
            f = getattr(lib, name)
            f.argtypes = [p, p, p, p, i64, p, p, p]
            f.restype = c_int
        for name in ("aucell_csr16_lut8_f32", "aucell_csr_lut8_f32"):
            f = getattr(lib, name)
            f.argtypes = [p, p, p, p, p, p, i64, p, p, p]
            f.restype = c_int
        lib.aucell_from_rankings_u32.argtypes = [p, p, i64, i64, p, p]
        lib.aucell_from_rankings_u32.restype = c_int
        lib.aucell_csr_f32_host.argtypes = [p, p, p, p, i64, p, p, i64]
        lib.aucell_csr_f32_host.restype = c_int
        lib.aucell_csr_is_canonical.argtypes = [p, p, i64, i64]
        lib.aucell_csr_is_canonical.restype = c_int
        lib.aucell_dense_strided_host.argtypes = [p, p, c_int, i64, i64, i64, p, p]
        lib.aucell_dense_strided_host.restype = c_int
        lib.aucell_cast_f64_to_f32.argtypes = [p, p, i64]
        lib.aucell_cast_f64_to_f32.restype = c_int
        lib.aucell_cast_i64_to_i32.argtypes = [p, p, i64]
        lib.aucell_cast_i64_to_i32.restype = c_int
        lib.aucell_host_dense_to_csr.argtypes = [p, c_int, i64, i64, i64, i64, p, p, p, i64, p]
        lib.aucell_host_dense_to_csr.restype = c_int
        lib.aucell_b200_release_staging.argtypes = [c_int]
        lib.aucell_b200_release_staging.restype = c_int
        lib.aucell_dense_f32_host.argtypes = [p, p, i64, i64, p, p, i64]
        lib.aucell_dense_f32_host.restype = c_int
        _lib = lib
        return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().aucell_b200_last_error()
        raise AUCellNativeError(rc, msg.decode("utf-8", "replace") if msg else "")


def device_info(device: int = 0):
    """(device_count, compute_capability, sm_count, smem_optin_bytes) -- raises without a usable device."""
    lib = load()
    n, cc, sms = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    smem = ctypes.c_int64(0)
    check(lib.aucell_b200_device_info(device, ctypes.byref(n), ctypes.byref(cc), ctypes.byref(sms), ctypes.byref(smem)))
    return n.value, cc.value, sms.value, smem.value


def launch_count() -> int:
    return int(load().aucell_b200_launch_count())
