import sys, torch
sys.path.insert(0, '.')
import bench
from dgod_b200.dg import DGFRCNN
dev = torch.device('cuda')
torch.manual_seed(0)
B = 2
model = DGFRCNN(9, B, "dg", bench.REG_WEIGHTS, 2).to(dev).train()
host = bench.synthetic_batches(2, B, 2, 0)
res = [bench.to_device(b, dev) for b in host]
bench.calibrate(model, res[0][0])
opt = model.configure_optimizer()
for it in range(4):
    imgs, boxes, labels, dom = res[it % 2]
    targets = [{"boxes": b.float(), "labels": l.long()} for b, l in zip(boxes, labels)]
    det = model.detector(imgs, targets)
    print(it, 'counts', [p.shape[0] for p in model.detector.last['proposals']])
    for d in det:
        print('   ', {k: float(v) for k, v in d['losses'].items()})
    f = model.detector.last['features']
    print('   feat absmax', {k: float(v.abs().max()) for k, v in f.items()})
    loss = sum(v for d in det for v in d["losses"].values())
    opt.zero_grad(); loss.backward()
    gn = torch.sqrt(sum((p.grad.float() ** 2).sum() for p in model.parameters() if p.grad is not None))
    print('   loss', float(loss), 'gradnorm', float(gn))
    opt.step()
    print('   nan params', sum(int(torch.isnan(p).any()) for p in model.parameters()))
