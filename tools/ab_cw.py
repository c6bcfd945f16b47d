"""Developer tool (GPU): A/B timing + bit fingerprints of column-walk kernel builds.

-DMIA_DEBUG libraries (the release library reads no environment variable), all from the same sources:
    libmia_scan_dbg.so    "new":  the defaults
    libmia_scan_old.so    "old":  -DMIA_CW_FMA=0 (h_t = a_t h_{t-1} + b_t as FMUL + FADD in the backward recompute; the default)
    libmia_scan_orig.so   "orig": -DMIA_CW_FMA=0 -DMIA_CW_OCT=0 (8-byte tile accesses: the first version of the kernels)
and, inside each, the planner knobs (MIA_CW_WIDE: the 255-register build of the backward; MIA_CW_STAGES; MIA_CW_MAXPERSM).

    python tools/ab_cw.py --build [--only old|orig]     # here (nvcc)
    python tools/ab_cw.py [--quick] --lib=KEY           # on the GPU box, ONE library per process: one JSON line per (shape, knobs)
                                                        # with fwd / bwd times and a fingerprint of every output tensor's bits
To ship the debug libraries to the GPU box take them out of .gpurunignore for that call (56 MB each).
Results of round 2: profiles/r2aa_ab_oct_wide.jsonl, profiles/r2ac_ab_fma_separate_processes.jsonl.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "medical_image_analysis_b200")
LIBS = {"new": os.path.join(PKG, "libmia_scan_dbg.so"), "old": os.path.join(PKG, "libmia_scan_old.so"),
        "orig": os.path.join(PKG, "libmia_scan_orig.so")}
OLD_FLAGS = {"old": ["-DMIA_CW_FMA=0"], "orig": ["-DMIA_CW_FMA=0", "-DMIA_CW_OCT=0"]}
KNOBS = ("MIA_CW_WIDE", "MIA_CW_STAGES", "MIA_CW_MAXPERSM")


def build():
    from medical_image_analysis_b200 import _build
    _build.build()
    nvcc = _build._nvcc()
    flags = [f for f in _build.NVCC_FLAGS if f not in ("-Xptxas", "-v")]
    csrc = os.path.join(PKG, "csrc")
    dbg = os.path.join(PKG, "build", "scan_api_dbg.o")
    subprocess.check_call([nvcc, *flags, "-DMIA_DEBUG", "-c", os.path.join(csrc, "scan_api.cu"), "-o", dbg])
    rel = {s: os.path.join(PKG, "build", s.replace(".cu", ".o")) for s in _build.SOURCES}
    subprocess.check_call([nvcc, "-shared", "-o", LIBS["new"], *[o for s, o in rel.items() if s != "scan_api.cu"], dbg,
                           "-gencode", "arch=compute_100a,code=sm_100a"])
    for key, extra in OLD_FLAGS.items():
        if "--only" in sys.argv and key not in sys.argv:
            continue
        os.makedirs(os.path.join(PKG, "build", key), exist_ok=True)
        procs, old = [], dict(rel)
        for s in _build.SOURCES:
            if s.startswith(("scan_fwd_", "scan_bwd_")):
                old[s] = os.path.join(PKG, "build", key, s.replace(".cu", ".o"))
                procs.append(subprocess.Popen([nvcc, *flags, *extra, "-c", os.path.join(csrc, s), "-o", old[s]]))
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("nvcc failed")
        subprocess.check_call([nvcc, "-shared", "-o", LIBS[key], *[o for s, o in old.items() if s != "scan_api.cu"], dbg,
                               "-gencode", "arch=compute_100a,code=sm_100a"])
    print(LIBS)


def use(lib_key):
    from medical_image_analysis_b200 import _lib
    _lib._lib = None
    _lib.LIB_PATH = LIBS[lib_key]


def main():
    import torch
    from medical_image_analysis_b200 import scan_bwd, scan_fwd
    quick = "--quick" in sys.argv
    D, G = 3072, 4
    # (B, L, dtype, out_f32, D, G)
    shapes = [(148, 196, torch.bfloat16, False, D, G), (16, 6400, torch.bfloat16, False, D, G)]
    if not quick:
        shapes += [(64, 196, torch.bfloat16, False, D, G), (32, 1024, torch.bfloat16, False, D, G), (64, 196, torch.bfloat16, True, D, G),
                   (3, 100, torch.float16, False, 256, 2), (3, 264, torch.bfloat16, False, 128, 2), (2, 196, torch.float32, False, 128, 2)]
    configs = [("old", {}), ("new", {}), ("old", {"MIA_CW_WIDE": "0"}), ("new", {"MIA_CW_WIDE": "0"})]
    for arg in sys.argv:                                               # --lib=old|new: one library per process (checksums compare runs)
        if arg.startswith("--lib="):
            configs = [(arg.split("=")[1], {}), (arg.split("=")[1], {"MIA_CW_WIDE": "0"})]
    reps = 10
    for B, L, dt, of32, Dm, Gm in shapes:
        g = torch.Generator(device="cuda").manual_seed(1)
        u = torch.randn(B, Dm, L, device="cuda", generator=g).to(dt)
        delta = (0.5 * torch.rand(B, Dm, L, device="cuda", generator=g)).to(dt)
        A = -0.5 * torch.rand(Dm, 1, device="cuda", generator=g)
        Bm = torch.randn(B, Gm, 1, L, device="cuda", generator=g).to(dt)
        Cm = torch.randn(B, Gm, 1, L, device="cuda", generator=g).to(dt)
        Dv = torch.randn(Dm, device="cuda", generator=g)
        bias = 0.5 * torch.rand(Dm, device="cuda", generator=g)
        dout = torch.randn(B, Dm, L, device="cuda", generator=g).to(torch.float32 if of32 else dt)
        base = None
        for lib_key, env in configs:
            for k in KNOBS:
                os.environ.pop(k, None)
            os.environ.update(env)
            use(lib_key)
            rec = {"B": B, "L": L, "dtype": str(dt).replace("torch.", ""), "of32": of32, "D": Dm, "lib": lib_key, "knobs": env}
            try:
                def fwd():
                    return scan_fwd(u, delta, A, Bm, Cm, Dv, None, bias, True, of32, want_block_states=True)
                out, x, _, hblk = fwd()

                def bwd():
                    return scan_bwd(u, delta, A, Bm, Cm, Dv, None, bias, dout, x, None, True, hblk=hblk)
                res = bwd()
                for fn, key in ((fwd, "fwd_ms"), (bwd, "bwd_ms")):
                    for _ in range(3):
                        fn()
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(reps):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    rec[key] = round(e0.elapsed_time(e1) / reps, 4)
                tensors = [out, x] + [t for t in res if t is not None]
                if base is None:
                    base = [t.clone() for t in tensors]
                    rec["bitwise_equal_to_first"] = None
                else:
                    rec["bitwise_equal_to_first"] = all(torch.equal(t, b) for t, b in zip(tensors, base))
                    if not rec["bitwise_equal_to_first"]:
                        rec["max_abs_diff"] = [float((t.float() - b.float()).abs().max()) for t, b in zip(tensors, base)]
                rec["hblk"] = hblk is not None
                # fingerprints of the raw bits (compare across processes / libraries)
                def fp(t):
                    b = t.contiguous().view(torch.uint8).to(torch.int64)
                    w = torch.arange(b.numel(), device=b.device) % 251 + 1
                    return int((b.flatten() * w).sum().item() % 2147483647)
                rec["bits"] = [fp(t) for t in tensors]
            except Exception as e:                                    # noqa: BLE001
                rec["error"] = str(e)[:300]
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    if "--build" in sys.argv:
        build()
    else:
        main()
