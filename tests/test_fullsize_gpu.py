"""GPU: BASELINE.json's full sizes, checked through size-independent properties (the CPU oracle
cannot cover 1e9 bins in seconds): Parseval per frame, frame/stream shift consistency, exact
power-of-two scaling, channel-permutation invariance, known-delay recovery."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_config2_full_size_parseval_and_shift():
    """256 ch x 4096 frames, N=2048, hop=1024 (config #2): sum((x*w)^2) == N*(P0 + 2*sum P_k + P_N/2)."""
    import torch
    from friture_b200 import audioproc
    C, F, N, H = 256, 4096, 2048, 1024
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randn((C, N + (F - 1) * H), generator=g, device="cuda") * 0.1
    p = audioproc()
    p.set_fftsize(N)
    P = p.stft(x, hop=H, log=False)
    assert tuple(P.shape) == (C, F, N // 2 + 1)
    w = torch.from_numpy(p.window.astype(np.float32)).cuda()
    worst = 0.0
    for c0 in range(0, C, 32):
        fr = x[c0:c0 + 32].unfold(1, N, H)                        # [32, F, N] view
        lhs = ((fr * w) ** 2).sum(-1, dtype=torch.float64)
        Pc = P[c0:c0 + 32].double()
        rhs = N * (Pc[..., 0] + 2 * Pc[..., 1:-1].sum(-1) + Pc[..., -1])
        worst = max(worst, float(((lhs - rhs).abs() / lhs).max()))
    assert worst < 2e-5, worst
    # log mode == 10 log10(power + 1e-30) of the power mode
    L = p.stft(x[:8], hop=H, log=True)
    ref = 10 * torch.log10(P[:8].double() + 1e-30)
    assert float((L.double() - ref).abs().max()) < 2e-4
    # a frame does not depend on where it sits in the batch / stream
    P2 = p.stft(x[5:6, 7 * H:], hop=H, log=False)
    assert torch.equal(P2[0, :100], P[5, 7:107])


def test_config3_full_size_scaling_and_permutation():
    """1024 ch x 64 blocks of 512: energies(2x) == 4 energies(x) exactly; channels independent."""
    import torch
    from friture_b200.octavefilters import Octave_Filters
    C, B, NB = 1024, 512, 64
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((C, B * NB), generator=g, device="cuda") * 0.1
    e1 = Octave_Filters(3).energies_batch(x, block=B)
    e2 = Octave_Filters(3).energies_batch(2 * x, block=B)
    assert tuple(e1.shape) == (C, NB, 27)
    assert torch.equal(e2, 4 * e1)
    perm = torch.randperm(C, generator=torch.Generator().manual_seed(1)).cuda()
    e3 = Octave_Filters(3).energies_batch(x[perm].contiguous(), block=B)
    assert torch.equal(e3, e1[perm])
    assert bool(torch.isfinite(e1).all()) and float(e1.min()) >= 0.0
    # a channel's numbers do not depend on its neighbours (same kernel variant: same bits) ...
    e4 = Octave_Filters(3).energies_batch(x[:8].contiguous(), block=B)
    assert torch.equal(e4, e1[:8])
    # ... and at 8192 channels, where the dispatcher picks 32-sample steps with two channels per
    # lane, only the rounding of the smoothing sums differs
    big = torch.cat([x[:8]] * 1024, dim=0)[:8192, :B * 4].contiguous()
    e5 = Octave_Filters(3).energies_batch(big, block=B)
    ref = e1[:8, :4]
    assert float(((e5[:8] - ref).abs() / ref.abs().amax(-1, keepdim=True)).max()) < 2e-6


def test_config4_full_size_delays():
    """4096 pairs x L=24000: every known integer delay is recovered."""
    import torch
    from friture_b200.correlation import GccPhat
    P, L = 4096, 24000
    g = torch.Generator(device="cuda").manual_seed(3)
    d0 = torch.randn((P, L), generator=g, device="cuda")
    k = torch.randint(0, L, (P,), generator=g, device="cuda")
    idxs = (torch.arange(L, device="cuda")[None, :] - k[:, None]) % L
    d1 = torch.gather(d0, 1, idxs) + 0.1 * torch.randn((P, L), generator=g, device="cuda")
    idx, val, _ = GccPhat(L).estimate(d0, d1, smooth=False)
    assert torch.equal(idx.long(), k)
    assert float(val.min()) > 0.1
