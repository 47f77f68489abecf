"""Build (nvcc, in-tree) and load libdimb200.so, the C-ABI declared in include/dimb200.h.

The library is the product: there is no Python/CPU fallback.  `load()` raises if the shared object is missing and cannot
be built, and every entry point returns DIM_ENODEVICE (raised here as RuntimeError) without an sm_100 GPU.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(_HERE, "libdimb200.so")
HEADER = os.path.join(ROOT, "include", "dimb200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

_lock = threading.Lock()
_lib = None


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libdimb200.so next to this file (cross-compiles without a GPU).
    Safe when several ranks start at once (torchrun): builds are serialised by a file lock, each writes its own temporary and
    publishes it with an atomic rename; the ranks that waited find a fresh library and skip the compile."""
    if not force and not _stale():
        return SO_PATH
    import fcntl
    with open(SO_PATH + ".lock", "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return SO_PATH
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            if not os.path.exists(nvcc):
                nvcc = "nvcc"
            # one object per source, compiled in parallel and only when stale (build/ is git-ignored), then one link
            from concurrent.futures import ThreadPoolExecutor
            objdir = os.path.join(_HERE, "build")
            os.makedirs(objdir, exist_ok=True)
            hdr_t = max(os.path.getmtime(d) for d in glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER])

            def compile_one(src):
                obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
                if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
                    return obj, None
                cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", src, "-o", obj]
                if verbose:
                    print(" ".join(cmd))
                r = subprocess.run(cmd, capture_output=True, text=True)
                return obj, (r.stdout + r.stderr if r.returncode != 0 else None)

            with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
                results = list(ex.map(compile_one, sources()))
            errs = [e for _, e in results if e]
            if errs:
                raise RuntimeError("nvcc failed:\n" + "\n".join(errs))
            tmp = f"{SO_PATH}.tmp.{os.getpid()}"
            cmd = [nvcc, "-shared", "-o", tmp] + [o for o, _ in results]
            if verbose:
                print(" ".join(cmd))
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.unlink(tmp)
                raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
            os.replace(tmp, SO_PATH)
            return SO_PATH
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)


class VQConfigC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("in_dim", "hidden", "layers", "heads", "ffn", "n_embed", "zdim", "pe_max_len")] \
        + [("neg_slope", C.c_float), ("fqn", C.c_int32), ("out_dim", C.c_int32)]


class S2SConfigC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("dim_in", "dim", "dim_audio", "depth", "heads", "dim_head", "max_seq_len",
                                         "num_tokens", "ff_mult")]


class ProfEntryC(C.Structure):
    _fields_ = [("category", C.c_int32), ("launches", C.c_int32), ("ms", C.c_double), ("bytes", C.c_double),
                ("flops", C.c_double)]


P, I, F, SZ, I64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_int64

# name -> (restype, argtypes); must list every function declared in include/dimb200.h (checked by tests/test_abi.py)
SIGNATURES = {
    "dim_last_error": (C.c_char_p, []),
    "dim_version": (I, []),
    "dim_launch_count": (C.c_uint64, []),
    "dim_profile_enable": (I, [I]),
    "dim_profile_collect": (I, [C.POINTER(ProfEntryC), I, C.POINTER(I)]),
    "dim_profile_category_name": (C.c_char_p, [I]),
    "dim_vq_argmin": (I, [P, P, P, I, I, I, P]),
    "dim_vq_gather": (I, [P, P, P, I, I, I, P, P]),
    "dim_linear_f32": (I, [P, I, P, P, P, I, P, I, I, I, I, I, F, P]),
    "dim_split_bf16_planes": (I, [P, I, I, I, I, P, P]),
    "dim_linear_bf16_planes": (I, [P, P, I, I, P, P, I, P, I, I, I, I, F, P]),
    "dim_conv5_leaky_f32": (I, [P, P, P, P, P, I, I, I, F, P]),
    "dim_repack_conv_weight": (I, [P, P, I, I, P]),
    "dim_instance_norm_f32": (I, [P, P, I, I, I, F, P]),
    "dim_layer_norm_f32": (I, [P, P, P, P, I, I, F, P]),
    "dim_linear_ragged_workspace_bytes": (SZ, [I, I, I]),
    "dim_linear_ragged_f32": (I, [P, I, P, I, P, P, I, I, I, I, I, F, P, SZ, P]),
    "dim_lstm_layer_workspace_bytes": (SZ, [I, I, I, I, I]),
    "dim_lstm_layer_f32": (I, [P, I, P, P, P, P, P, P, P, P, I, I, I, P, P, SZ, P]),
    "dim_resample_features": (I, [P, I, I, I, I, I, P, P]),
    "dim_assemble_batch": (I, [P, P, P, P, I, I, I, I, P, P, P, P]),
    "dim_attention_f32": (I, [P, I, P, I, P, I, P, I, P, P, I, I, I, I, I, F, I, P]),
    "dim_create": (I, [C.POINTER(P), I]),
    "dim_destroy": (I, [P]),
    "dim_set_tensor": (I, [P, C.c_char_p, P, I, I, C.POINTER(I64)]),
    "dim_vqvae_build": (I, [P, C.c_char_p, C.POINTER(VQConfigC), I, C.POINTER(I)]),
    "dim_vqvae_build_parts": (I, [P, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(VQConfigC), I, C.POINTER(I)]),
    "dim_vqvae_workspace_bytes": (SZ, [P, I, I, I]),
    "dim_vqvae_encode": (I, [P, I, P, P, P, I, I, P, P, P, P, SZ, P]),
    "dim_vqvae_decode": (I, [P, I, P, P, P, I, I, P, P, SZ, P]),
    "dim_slmft_build": (I, [P, C.POINTER(S2SConfigC), I, C.POINTER(I)]),
    "dim_slmft_workspace_bytes": (SZ, [P, I, I, I, I]),
    "dim_slmft_context": (I, [P, I, P, P, P, I, I, P, P, P, SZ, P]),
    "dim_slmft_encode": (I, [P, I, I, P, P, P, I, I, I, I, P, P, SZ, P]),
    "dim_slmft_generate": (I, [P, I, P, P, P, I, I, I, F, I, P, P, P, P, SZ, P]),
    "dim_slmft_samples_workspace_bytes": (SZ, [P, I, I, I, I, I]),
    "dim_slmft_generate_samples": (I, [P, I, P, P, P, I, I, I, I, F, I, P, P, P, P, SZ, P]),
    "dim_slmft_teacher_forced_workspace_bytes": (SZ, [P, I, I, I, I]),
    "dim_slmft_teacher_forced": (I, [P, I, P, P, P, P, I, I, I, P, P, SZ, P]),
    "dim_decode_trace_enable": (I, [I]),
    "dim_decode_trace_collect": (I, [C.POINTER(C.c_double), C.POINTER(C.c_int32), I, C.POINTER(I)]),
    "dim_decode_set_impl": (I, [I]),
}


def load(build_if_missing: bool = True):
    """Return the ctypes handle of libdimb200.so (building it first when the sources are newer)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if build_if_missing:
            try:
                build()
            except Exception:
                if not os.path.exists(SO_PATH):
                    raise
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                               "there is no fallback path")
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
        return lib


def check(code: int, what: str = ""):
    if code != 0:
        msg = load().dim_last_error()
        raise RuntimeError(f"libdimb200 {what} failed (code {code}): {msg.decode() if msg else ''}")


def profile_enable(on: bool):
    check(load().dim_profile_enable(int(on)), "dim_profile_enable")


def profile_collect():
    """-> list of dict(category, launches, ms, bytes, flops) since the last collect (synchronises the device)."""
    lib = load()
    arr = (ProfEntryC * 32)()
    n = C.c_int(0)
    check(lib.dim_profile_collect(arr, 32, C.byref(n)), "dim_profile_collect")
    return [dict(category=lib.dim_profile_category_name(arr[i].category).decode(), launches=arr[i].launches,
                 ms=arr[i].ms, bytes=arr[i].bytes, flops=arr[i].flops) for i in range(n.value)]


MK_TYPE_NAMES = {1: "gemm", 2: "attention", 3: "residual_layernorm", 4: "gelu", 5: "sample_embed_layernorm", 6: "nop"}


def decode_trace_enable(on: bool):
    check(load().dim_decode_trace_enable(int(on)), "dim_decode_trace_enable")


def decode_trace_collect():
    """-> list of (type name, ms) per phase of one decode step, summed over the steps of the last generate call."""
    lib = load()
    ms = (C.c_double * 64)()
    ty = (C.c_int32 * 64)()
    n = C.c_int(0)
    check(lib.dim_decode_trace_collect(ms, ty, 64, C.byref(n)), "dim_decode_trace_collect")
    return [(MK_TYPE_NAMES.get(ty[i], str(ty[i])), ms[i]) for i in range(n.value)]


def decode_set_impl(impl: int):
    check(load().dim_decode_set_impl(int(impl)), "dim_decode_set_impl")
