"""TEST INFRASTRUCTURE: numpy (uint64, (...,2)) <-> torch CUDA (int64, (...,2)) for quad bit patterns."""
import numpy as np
import torch


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def to_host(t):
    return t.detach().cpu().numpy().view(np.uint64)


def dev_random(shape, kind="D113", seed=0, device="cuda"):
    """Device-side generation of quads (int64 (...,2)); same distributions as quad.random_quads:
    D53 = doubles U(-1,1) cast exactly, D113 = full random mantissa with U(-1,1)-like exponents,
    Dexp = D113 x 2^U{-40..40}."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    shape = tuple(shape)
    if kind == "D53":
        d = torch.rand(shape, generator=g, device=device, dtype=torch.float64) * 2 - 1
        b = d.view(torch.int64)
        sign = b & (-(1 << 63))
        e = (b >> 52) & 0x7FF
        m = b & ((1 << 52) - 1)
        hi = sign | ((e - 1023 + 16383) << 48) | (m >> 4)
        hi = torch.where(d == 0, sign, hi)
        lo = m << 60
        return torch.stack([lo, hi], dim=-1)
    lo = torch.randint(-(1 << 63), (1 << 63) - 1, shape, generator=g, device=device, dtype=torch.int64)
    mh = torch.randint(0, 1 << 48, shape, generator=g, device=device, dtype=torch.int64)
    s = torch.randint(0, 2, shape, generator=g, device=device, dtype=torch.int64)
    # exponent of a U(0,1) variate: -1 - Geometric(1/2), truncated at -40
    u = torch.rand(shape, generator=g, device=device, dtype=torch.float64).clamp_min(2.0 ** -40)
    e = torch.floor(torch.log2(u)).to(torch.int64)
    if kind == "Dexp":
        e = e + torch.randint(-40, 41, shape, generator=g, device=device, dtype=torch.int64)
    hi = (s << 63) | ((e + 16383) << 48) | mh
    return torch.stack([lo, hi], dim=-1)
