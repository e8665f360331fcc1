"""Development aid: build gemm_tc.cu variants (-D flags) into tools/_variants/ and time the four
per-layer GEMM shapes of cfg2 through the C ABI (cse_linear).  Not part of the product or the bench.

  python tools/gemm_variants.py build            # here (nvcc cross-compiles)
  python tools/gemm_variants.py run [name ...]   # on the GPU box
"""
import ctypes as C
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "contextual-speech-extraction_b200", "csrc")
OUT = os.path.join(ROOT, "tools", "_variants")
NVCC = "/usr/local/cuda/bin/nvcc"
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden"]

BASE = ["-DCSE_EPI_WARPS=8", "-DCSE_STAGES_256=4"]
VARIANTS = {
    "r128": BASE + ["-DCSE_EPI_ROW_BYTES=128"],
    "r64": BASE + ["-DCSE_EPI_ROW_BYTES=64"],
    "r64_nostore": BASE + ["-DCSE_EPI_ROW_BYTES=64", "-DCSE_DBG_NOSTORE"],
    "r64_noepi": BASE + ["-DCSE_EPI_ROW_BYTES=64", "-DCSE_DBG_NOEPI"],
    "ffn_g1only": BASE + ["-DFFN_DBG_NOG2", "-DFFN_DBG_NOE1", "-DFFN_DBG_NOFINAL"],
    "ffn_g2only": BASE + ["-DFFN_DBG_NOG1", "-DFFN_DBG_NOE1", "-DFFN_DBG_NOFINAL"],
    "ffn_e1spin": BASE + ["-DFFN_E1_SPIN"],
    "ffn_e1spin_noe1_nofinal": BASE + ["-DFFN_E1_SPIN", "-DFFN_DBG_NOE1", "-DFFN_DBG_NOFINAL"],
    "ffn_g1n64": BASE + ["-DFFN_DBG_G1N=64", "-DFFN_DBG_NOE1", "-DFFN_DBG_NOFINAL"],
    "ffn_g2n128": BASE + ["-DFFN_DBG_G2N=128", "-DFFN_DBG_NOE1", "-DFFN_DBG_NOFINAL"],
    "ffn_g1n64_g2n128": BASE + ["-DFFN_DBG_G1N=64", "-DFFN_DBG_G2N=128", "-DFFN_DBG_NOE1", "-DFFN_DBG_NOFINAL"],
    "ffn_solo": BASE + ["-DFFN_SOLO"],
    "ffn_solo_noe1_nofinal": BASE + ["-DFFN_SOLO", "-DFFN_DBG_NOE1", "-DFFN_DBG_NOFINAL"],
    "ffn_solo_loadsonly": BASE + ["-DFFN_SOLO", "-DFFN_DBG_NOE1", "-DFFN_DBG_NOFINAL", "-DFFN_DBG_NOMMA"],
    "ffn_g1ts": BASE + ["-DFFN_DBG_G1TS", "-DFFN_DBG_NOE1", "-DFFN_DBG_NOFINAL"],
    "ffn_g1ts_g1only": BASE + ["-DFFN_DBG_G1TS", "-DFFN_DBG_NOG2", "-DFFN_DBG_NOE1", "-DFFN_DBG_NOFINAL"],
    "ffn_trace": BASE + ["-DFFN_TRACE"],
    "ffn_trace_noe1_nofinal": BASE + ["-DFFN_TRACE", "-DFFN_DBG_NOE1", "-DFFN_DBG_NOFINAL"],
    "ffn_stages3": BASE + ["-DFFN_STAGES=3"],
    "ffn_nofinal": BASE + ["-DFFN_DBG_NOFINAL"],
    "ffn_noe1": BASE + ["-DFFN_DBG_NOE1"],
    "ffn_loadsonly": BASE + ["-DFFN_DBG_NOE1", "-DFFN_DBG_NOFINAL", "-DFFN_DBG_NOMMA"],
    "ffn_nomma": BASE + ["-DFFN_DBG_NOMMA"],
    "ffn_noe1_nofinal": BASE + ["-DFFN_DBG_NOE1", "-DFFN_DBG_NOFINAL"],
    "r128_loadsonly": BASE + ["-DCSE_EPI_ROW_BYTES=128", "-DCSE_DBG_NOEPI", "-DCSE_DBG_NOMMA"],
    "r128_nomma": BASE + ["-DCSE_EPI_ROW_BYTES=128", "-DCSE_DBG_NOMMA"],
    "r64_nomma": BASE + ["-DCSE_EPI_ROW_BYTES=64", "-DCSE_DBG_NOMMA"],
}


def build(names):
    os.makedirs(OUT, exist_ok=True)
    varied = ["gemm_tc", "ffn_tc"]
    others = [o for o in glob.glob(os.path.join(CSRC, "*.o")) if os.path.basename(o)[:-2] not in varied]
    for name in names or VARIANTS:
        objs = []
        for src in varied:
            obj = os.path.join(OUT, f"{src}_{name}.o")
            subprocess.check_call([NVCC] + FLAGS + VARIANTS[name] + ["-c", os.path.join(CSRC, src + ".cu"), "-o", obj])
            objs.append(obj)
        lib = os.path.join(OUT, f"libcse_{name}.so")
        subprocess.check_call([NVCC, "-shared", "-o", lib] + objs + others + ["-gencode", "arch=compute_100a,code=sm_100a"])
        print("built", lib)


def run(names):
    import torch
    M = 136544
    shapes = [("qkv ", 768, 256, 0, 0), ("outp", 256, 256, 0, 1), ("ffn1", 1024, 256, 1, 0), ("ffn2", 256, 1024, 0, 1)]
    torch.manual_seed(0)
    for name in names or sorted(n[7:-3] for n in os.listdir(OUT) if n.startswith('libcse_')):
        lib = C.CDLL(os.path.join(OUT, f"libcse_{name}.so"))
        lib.cse_linear.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        lib.cse_last_error.restype = C.c_char_p
        line = [f"{name:12s}"]
        for tag, N, K, relu, resid in (shapes if not name.startswith("ffn_") else []):
            A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
            W = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
            b = torch.randn(N, device="cuda")
            out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if resid else torch.bfloat16)
            st = torch.cuda.current_stream().cuda_stream

            def call(bias=True):
                rc = lib.cse_linear(A.data_ptr(), K, W.data_ptr(), b.data_ptr() if bias else None, 1.0,
                                    out.data_ptr() if resid else None, out.data_ptr(), N, M, N, K, relu, resid, 1, st)
                if rc:
                    raise RuntimeError(lib.cse_last_error().decode())
            res = []
            for bias in (True, False):
                for _ in range(3):
                    call(bias)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    call(bias)
                e1.record()
                torch.cuda.synchronize()
                res.append(e0.elapsed_time(e1) / 20 * 1e3)
            line.append(f"{tag} {res[0]:6.1f}/{res[1]:6.1f}")
        if hasattr(lib, "cse_ffn_fused"):
            lib.cse_ffn_fused.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_void_p]
            A = (torch.randn(M, 256, device="cuda") * 0.5).bfloat16()
            W1 = (torch.randn(1024, 256, device="cuda") * 0.05).bfloat16()
            W2 = (torch.randn(256, 1024, device="cuda") * 0.03).bfloat16()
            b1, b2 = torch.randn(1024, device="cuda"), torch.randn(256, device="cuda")
            R = torch.zeros(M, 256, device="cuda")
            st = torch.cuda.current_stream().cuda_stream

            def ffn():
                if lib.cse_ffn_fused(A.data_ptr(), W1.data_ptr(), b1.data_ptr(), W2.data_ptr(), b2.data_ptr(),
                                     R.data_ptr(), M, st):
                    raise RuntimeError(lib.cse_last_error().decode())
            for _ in range(3):
                ffn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ffn()
            e1.record()
            torch.cuda.synchronize()
            line.append(f"ffn_fused {e0.elapsed_time(e1) / 20 * 1e3:6.1f}")
            if hasattr(lib, "cse_debug_ffn_trace"):
                buf = (C.c_ulonglong * (4 * 64))()
                lib.cse_debug_ffn_trace(buf)
                t0 = buf[0]
                t0 = buf[0]
                print("MMA thread, block 0: per ring stage [wait_start, deps_ok, weights_ok] in cycles; 8 stages per chunk (4 G1 + 4 G2), 4 chunks per tile")
                for k in range(0, 64):
                    print(f"  tile {k // 32} chunk {(k // 8) % 4} {'G1' if k % 8 < 4 else 'G2'}{k % 4}"
                          f" {buf[k] - t0:8d} {buf[64 + k] - t0:8d}")
                print("E1 warp 2, tile 1 chunk 1: hfull seen", int(buf[192 + 39] - t0), "; per piece [ld done, math done, bar done, st done, arrived]")
                for pp in range(4):
                    print("   piece", pp, [int(buf[192 + 40 + pp * 5 + j] - t0) for j in range(5)])
                print("output warp 10, block 0: yfull seen, then after each of the 8 tmem loads")
                for it in range(2):
                    print("  tile", it, [int(buf[128 + it * 16 + j] - t0) for j in range(9)])
        print("  ".join(line) + "   (us with bias / without)", flush=True)


if __name__ == "__main__":
    (build if sys.argv[1] == "build" else run)(sys.argv[2:])
