"""Seeded synthetic WSSS4LUAD / BCSS-shaped inputs for the parity tests and bench.py.

Everything is generated on the CPU with ``torch.Generator().manual_seed`` so that the CPU oracle and the GPU
library see identical bits; nothing here reads a dataset or the reference tree.
"""
import torch


def view_sizes(T, scales, stride=8):
    """Low-resolution side of each (scale, flip) view: floor(T*s/stride), two views (no-flip, hflip) per scale."""
    out = []
    for s in scales:
        h = int(T * s) // stride
        out += [h, h]
    return out


def make_views(N, C, sizes, seed, std=3.0, flips=None):
    """[N,C,h,h] float32 logits ~ N(0, std^2) per view (seed + v) and the de-augmentation codes (hflip for odd v)."""
    views = []
    for v, h in enumerate(sizes):
        g = torch.Generator().manual_seed(seed + v)
        views.append(torch.randn((N, C, h, h), generator=g) * std)
    if flips is None:
        flips = [4 * (v % 2) for v in range(len(sizes))]
    return views, flips


def make_gt(N, T, C, seed, ignore_frac=0.15, ignore_label=None, block=16):
    """Blocky ground truth: (T/block)^2 uniform labels in [0,C) nearest-upsampled, ``ignore_frac`` of blocks set to
    ``ignore_label`` (default C: 3 = background for WSSS4LUAD, 4 = white for BCSS)."""
    g = torch.Generator().manual_seed(seed)
    nb = (T + block - 1) // block
    lab = torch.randint(0, C, (N, nb, nb), generator=g, dtype=torch.uint8)
    ign = torch.rand((N, nb, nb), generator=g) < ignore_frac
    lab[ign] = C if ignore_label is None else ignore_label
    gt = lab.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :T, :T].contiguous()
    return gt


def make_present(N, C, seed, single_frac=0.4):
    """Image-level label vectors: ``single_frac`` of tiles have exactly one class, the rest at least two."""
    g = torch.Generator().manual_seed(seed)
    p = torch.zeros((N, C), dtype=torch.uint8)
    single = torch.rand(N, generator=g) < single_frac
    first = torch.randint(0, C, (N,), generator=g)
    extra = torch.rand((N, C), generator=g) < 0.5
    for n in range(N):
        p[n, first[n]] = 1
        if not single[n]:
            p[n] |= extra[n].to(torch.uint8)
            if int(p[n].sum()) < 2:
                p[n, (int(first[n]) + 1) % C] = 1
    return p


def cfg1(N=256, T=224, C=3):
    """BASELINE config 1: single stride-8 view, bg + gt + labels + confusion."""
    views, codes = make_views(N, C, [T // 8], 1001, flips=[0])
    gt = make_gt(N, T, C, 1002)
    return dict(views=views, codes=codes, T=T, C=C, gt=gt, bg=(gt == C).to(torch.uint8), present=None)


def cfg2(N=1024, T=224, C=3, scales=(0.75, 1.0, 1.25), single_frac=0.4):
    """BASELINE config 2: pseudo-mask inference, V=6 (3 scales x flip), present vector, bg, 32x32 logits export."""
    views, codes = make_views(N, C, view_sizes(T, scales), 2001)
    gt = make_gt(N, T, C, 2002)
    return dict(views=views, codes=codes, T=T, C=C, gt=None, bg=(gt == C).to(torch.uint8), present=make_present(N, C, 2003, single_frac))


def cfg3(N=1000, T=224, C=4, scales=(0.75, 1.0, 1.25)):
    """BASELINE config 3: BCSS-shaped, V=6, gt in {0..4} (4 ignored, 5 %), confusion, no bg / present."""
    views, codes = make_views(N, C, view_sizes(T, scales), 3001)
    gt = make_gt(N, T, C, 3002, ignore_frac=0.05)
    return dict(views=views, codes=codes, T=T, C=C, gt=gt, bg=None, present=None)


def cfg5(N=8, T=512, C=4, scales=(1, 1.25, 1.5, 1.75, 2)):
    """BASELINE config 5: large tiles, V=10."""
    views, codes = make_views(N, C, view_sizes(T, scales), 5001)
    gt = make_gt(N, T, C, 5002, ignore_frac=0.05)
    return dict(views=views, codes=codes, T=T, C=C, gt=gt, bg=None, present=None)


def family_views(name, N, sizes, gen, dev, C=3):
    """Input families for the data-dependence record of the filtered kernels (tests/test_gpu_filter.py, tools/bench_all.py): seeded view
    sets generated ON the device, [N,C,h,h] per view.  gauss: i.i.d. N(0, 3^2) (the bench default); smooth: low-frequency fields (large
    uniform regions, long class boundaries -- what a trained backbone emits); neartie / quantized / extreme: adversarial."""
    views = []
    for h in sizes:
        if name == "gauss":
            v = torch.randn((N, C, h, h), generator=gen, device=dev) * 3
        elif name == "smooth":          # low-frequency fields + a little noise: large uniform regions, long class boundaries
            coarse = torch.randn((N, C, 4, 4), generator=gen, device=dev) * 3
            v = torch.nn.functional.interpolate(coarse, (h, h), mode="bilinear", align_corners=True) + torch.randn((N, C, h, h), generator=gen, device=dev) * 0.05
        elif name == "neartie":         # classes differ by 1e-7 .. 1e-2 almost everywhere
            base = torch.randn((N, 1, h, h), generator=gen, device=dev) * 3
            eps = 10 ** (torch.rand((N, 1, 1, 1), generator=gen, device=dev) * 5 - 7)
            v = base + torch.randn((N, C, h, h), generator=gen, device=dev) * eps
        elif name == "quantized":       # multiples of 0.5: exact ties everywhere
            v = torch.round(torch.randn((N, C, h, h), generator=gen, device=dev) * 2) / 2
        elif name == "extreme":         # denormals, tiny and huge magnitudes per tile
            mag = 10 ** (torch.rand((N, 1, 1, 1), generator=gen, device=dev) * 58 - 40)
            v = torch.randn((N, C, h, h), generator=gen, device=dev) * mag
        else:
            raise KeyError(name)
        views.append(v.contiguous())
    return views
