"""Round-1 GPU-box probe: how torch-CUDA relates to torch-CPU on this path's ops.
Writes gpurun_out/probe_torch.json.  Not product code."""
import json, os, sys, time
import torch
import torch.nn.functional as F

out = {"torch": torch.__version__, "cpu_count": os.cpu_count()}
p = torch.cuda.get_device_properties(0)
out["gpu"] = {"name": p.name, "sms": p.multi_processor_count, "mem_gb": p.total_memory / 2**30}
g = torch.Generator().manual_seed(7)

def cmp(a, b):
    a = a.double(); b = b.double()
    d = (a - b).abs()
    return {"bit_equal": bool(torch.equal(a, b)), "n_diff": int((d > 0).sum()), "max_abs": float(d.max()),
            "max_rel_gate": float((d / b.abs().clamp_min(1)).max())}

cases = [((16, 3, 28, 28), (224, 224), torch.float32), ((16, 3, 21, 21), (224, 224), torch.float32),
         ((16, 3, 35, 35), (224, 224), torch.float32), ((4, 3, 224, 224), (200, 180), torch.float32),
         ((4, 3, 224, 224), (32, 32), torch.float32), ((2, 4, 64, 64), (512, 512), torch.float32),
         ((1, 3, 300, 280), (240, 224), torch.float64), ((1, 3, 150, 130), (300, 260), torch.float64)]
res = []
for shp, size, dt in cases:
    x = (torch.randn(shp, generator=g) * 3).to(dt)
    c = F.interpolate(x, size, mode="bilinear")
    d = F.interpolate(x.cuda(), size, mode="bilinear").cpu()
    r = cmp(d, c); r["case"] = f"{shp}->{size} {dt}"; res.append(r)
out["interpolate_cuda_vs_cpu"] = res

x = torch.randn((8, 3, 224, 224), generator=g) * 3
sc, sg = torch.softmax(x, 1), torch.softmax(x.cuda(), 1).cpu()
out["softmax_cuda_vs_cpu"] = cmp(sg, sc)
out["argmax_softmax_mismatch"] = int((sc.argmax(1) != sg.argmax(1)).sum())
out["argmax_logit_vs_softmax_cpu"] = int((x.argmax(1) != sc.argmax(1)).sum())
xs = torch.randn((8, 3, 224, 224), generator=g) * 0.01
out["argmax_logit_vs_softmax_cpu_small"] = int((xs.argmax(1) != torch.softmax(xs, 1).argmax(1)).sum())
out["argmax_logit_vs_softmax_cuda_small"] = int((xs.cuda().argmax(1) != torch.softmax(xs.cuda(), 1).argmax(1)).sum().cpu())
t = torch.tensor([[1.0, 1.0, 0.5], [0.5, 2.0, 2.0], [3.0, 3.0, 3.0]]).T.contiguous()  # [C=3, 3 px]
out["argmax_tie_cuda"] = t.cuda().argmax(0).cpu().tolist()
out["argmax_tie_cpu"] = t.argmax(0).tolist()
xe = torch.linspace(-20, 0, 1 << 20)
ec, eg = torch.exp(xe), torch.exp(xe.cuda()).cpu()
out["exp_cuda_vs_cpu"] = cmp(eg, ec)

# torch-eager composite of cfg2 on this GPU (the "bar to beat"), rough timing
N, C, T = 1024, 3, 224
views = [(torch.randn((N, C, h, h), generator=g) * 3).cuda() for h in (21, 21, 28, 28, 35, 35)]
bg = (torch.rand((N, T, T), generator=g) < 0.15).cuda()
def eager():
    acc = None
    for i, v in enumerate(views):
        if i % 2: v = v.flip(3)
        u = F.interpolate(v, (T, T), mode="bilinear")
        acc = u if acc is None else acc + u
    fused = acc / len(views)
    low = fused[..., 3::7, 3::7].contiguous()
    lab = torch.softmax(fused, 1).argmax(1).to(torch.uint8)
    lab[bg] = C
    return lab, low
for _ in range(3): eager()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): eager()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
out["torch_eager_cfg2_tiles_per_s"] = N / dt
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe_torch.json", "w"), indent=1)
print(json.dumps(out, indent=1))
