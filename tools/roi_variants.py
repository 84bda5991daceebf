"""Experiment harness: build variants of roi_align_tma.cu with extra -D flags (CPU box), time them on the GPU.

  python tools/roi_variants.py build tagA:-DX=1,-DY tagB:...     -> tools/_variants/lib_<tag>.so
  python tools/roi_variants.py run [tags...]                     -> kernel_us of fwd/bwd per variant + max rel err vs 'base'
"""
import sys, subprocess, os, ctypes as C, statistics
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
VAR = ROOT / "tools" / "_variants"

def build(specs):
    from dgod_b200 import build as B
    B.build()
    VAR.mkdir(parents=True, exist_ok=True)
    procs = []
    for spec in specs:
        tag, _, flags = spec.partition(":")
        flags = [f for f in flags.split(",") if f]
        src_name = os.environ.get("VARIANT_SRC", "roi_align_tma.cu")     # which .cu the -D flags apply to
        src = B.CSRC / src_name
        if "@" in tag:
            tag, rev = tag.split("@")
            src = B.CSRC / f"_ref_{tag}.cu"
            src.write_text(subprocess.check_output(["git", "show", f"{rev}:dgod_b200/csrc/roi_align_tma.cu"], text=True))
        obj = VAR / f"{Path(src_name).stem}_{tag}.o"
        cmd = [B._nvcc(), *B.NVCC_FLAGS, *flags, "-c", str(src), "-o", str(obj)]
        procs.append((tag, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for tag, obj, p in procs:
        out, err = p.communicate()
        if p.returncode:
            print(tag, "FAILED\n", err[-3000:]); continue
        regs = [l for l in err.splitlines() if "Used" in l]
        names = [l for l in err.splitlines() if "Function properties" in l]
        for n, r in zip(names, regs):
            if "IfLi256ELi2" in n and "tma" in n: print(tag, n.split("_ZN4dgod")[1][:24], r.split(":")[1].strip()[:60])
        objs = [str(obj if s == src_name else B.OBJ_DIR / (Path(s).stem + ".o")) for s in B.SOURCES]
        subprocess.check_call([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(VAR / f"lib_{tag}.so"), *objs])
        print("built", tag)

def run_one(tag, dtype_name="f32", per=512, iters=15):
    import torch
    from dgod_b200 import _lib
    _lib.LIB_PATH = VAR / f"lib_{tag}.so" if tag != "tree" else _lib.LIB_PATH
    from dgod_b200 import ops, synth
    DEV = torch.device("cuda")
    dtype = torch.float32 if dtype_name == "f32" else torch.bfloat16
    B_, Cc, H, W = int(os.environ.get('NB', 8)), 256, 608, 1024
    per = int(os.environ.get('PER', per))
    feats = [f.to(DEV).contiguous(memory_format=torch.channels_last) for f in synth.random_features(B_, Cc, H, W, 0, dtype=dtype)]
    boxes = [synth.random_boxes(per, H, W, synth.gen(10 + i)) for i in range(B_)]
    rois = synth.rois_from_boxes(boxes).to(DEV)
    lib = _lib.load()
    K = rois.shape[0]
    go = torch.randn(K, Cc, 7, 7, generator=synth.gen(99)).to(DEV, dtype)
    out = torch.empty(K, Cc, 7, 7, device=DEV, dtype=dtype)
    grads = [torch.empty_like(f) for f in feats]
    cfg, _ = ops._roi_config(feats, [1/4, 1/8, 1/16, 1/32], 7, 7, 2, False, 2, 5, 224.0, 4.0)
    cfg.channels_last = 1
    wsb = max(lib.dgod_msroi_align_bwd_workspace_bytes_cfg(C.byref(cfg), K), lib.dgod_msroi_align_fwd_workspace_bytes(K))
    ws = torch.zeros(wsb, dtype=torch.uint8, device=DEV)
    fptrs, keep1 = ops._level_ptrs(feats)
    gptrs, keep2 = ops._level_ptrs(grads)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    st = ops._stream()
    def fwd(): ops.check(lib.dgod_msroi_align_fwd(C.byref(cfg), fptrs, ops._p(rois), K, ops._p(out), ops._p(ws), wsb, st))
    def bwd(): ops.check(lib.dgod_msroi_align_bwd(C.byref(cfg), ops._p(go), ops._p(rois), K, None, gptrs, int(os.environ.get('BWD_ALGO', 3)), ops._p(ws), wsb, st))
    res = {}
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        for _ in range(3): fn()
        ts = []
        for _ in range(iters):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res[name] = statistics.median(ts)
    if os.environ.get("TIMING"):
        ws[2048:2048 + 64].zero_(); bwd(); torch.cuda.synchronize()
        t = ws[2048:2048 + 64].view(torch.int64).cpu().tolist()
        names = ["wait plan", "wait G", "gr->regs", "T", "wait buf", "R", "fence+arrive", "total"]
        print("   " + "  ".join(f"{n}={v / 296 / 1965:.1f}us" for n, v in zip(names, t)))
    VAR.mkdir(parents=True, exist_ok=True)
    torch.save({"out": out.float().cpu(), "g": [g.float().cpu() for g in grads]}, VAR / f"res_{tag}.pt")
    print(f"{tag:14s} {dtype_name} fwd {res['fwd']:7.1f} us  bwd {res['bwd']:7.1f} us", flush=True)

def compare(tags):
    import torch
    base = torch.load(VAR / f"res_{tags[0]}.pt")
    for t in tags[1:]:
        r = torch.load(VAR / f"res_{t}.pt")
        eo = ((r["out"] - base["out"]).abs().max() / base["out"].abs().max()).item()
        eg = max(((a - b).abs().max() / b.abs().max()).item() for a, b in zip(r["g"], base["g"]))
        print(f"{t:14s} vs {tags[0]}: out relerr {eo:.2e}  grad relerr {eg:.2e}")

if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "one":
        run_one(*sys.argv[2:])
    else:
        tags = sys.argv[2:] or sorted(p.stem[4:] for p in VAR.glob("lib_*.so"))
        dt = os.environ.get("DT", "f32")
        for t in tags:
            subprocess.call([sys.executable, __file__, "one", t, dt])
        compare(tags)
