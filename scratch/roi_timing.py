"""Debug build only (-DDGOD_ROI_TIMING): prints where consumer thread 0 of the RoIAlign backward spends its time."""
import sys, ctypes as C
sys.path.insert(0, '.')
import torch
from dgod_b200 import ops, synth, _lib
DEV = torch.device("cuda")
B, Cc, H, W, per = 8, 256, 608, 1024, 512
feats = [f.to(DEV).contiguous(memory_format=torch.channels_last) for f in synth.random_features(B, Cc, H, W, 0)]
boxes = [synth.random_boxes(per, H, W, synth.gen(10 + i)) for i in range(B)]
rois = synth.rois_from_boxes(boxes).to(DEV)
lib = _lib.load()
K = rois.shape[0]
go = torch.randn(K, Cc, 7, 7, device=DEV)
grads = [torch.empty_like(f) for f in feats]
cfg, _ = ops._roi_config(grads, [1/4, 1/8, 1/16, 1/32], 7, 7, 2, False, 2, 5, 224.0, 4.0)
cfg.channels_last = 1
wsb = lib.dgod_msroi_align_bwd_workspace_bytes(K)
ws = torch.zeros(wsb, dtype=torch.uint8, device=DEV)
ptrs, keep = ops._level_ptrs(grads)
for it in range(3):
    ops.check(lib.dgod_msroi_align_bwd(C.byref(cfg), ops._p(go), ops._p(rois), K, None, ptrs, 3, ops._p(ws), wsb, ops._stream()))
    torch.cuda.synchronize()
t = ws[2048:2048 + 72].view(torch.int64).cpu().tolist()
names = ["wait plan", "wait G", "gr load + zero wait", "row compute", "wait_read", "barrier", "issue", "fence.proxy", "total"]
n_cta = 296
for n, v in zip(names, t):
    print(f"{n:22s} {v / n_cta / 1e3:9.1f} us per CTA")
