"""GPU parity of the batched fec::BurstInterleaver::deinterleave (src/fec/burst_interleaver.cpp:39-78)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("group", [1, 2, 3, 4, 8])
def test_burst_deinterleave_matches_reference(ctx, ref, group):
    from ria_b200 import fec
    rng = np.random.default_rng(group)
    n_groups = 5
    phys = rng.standard_normal((n_groups, group, 2600)).astype(np.float32)     # rows longer than 2592: the tail is ignored
    got = fec.burst_deinterleave_batch(torch.from_numpy(phys).cuda(), ctx).cpu().numpy()
    for g in range(n_groups):
        want = ref.burst_deinterleave(phys[g, :, :2592])
        assert np.array_equal(got[g].view(np.uint32), want.view(np.uint32)), (group, g)


def test_burst_roundtrip_against_tx_permutation(ctx):
    """interleave (:8-37, restated in numpy) then the device de-interleaver is the identity on bytes' soft bits."""
    from ria_b200 import fec
    N, B = 4, 324
    logical = np.arange(N * B, dtype=np.float32).reshape(N, B)
    physical = np.zeros_like(logical)
    for f in range(N):
        for b in range(B):
            flat = N * b + f
            physical[flat // B, flat % B] = logical[f, b]
    soft = np.repeat(physical, 8, axis=1)[None]                # every byte -> 8 equal soft bits
    got = fec.burst_deinterleave_batch(torch.from_numpy(np.ascontiguousarray(soft)).cuda(), ctx).cpu().numpy()[0]
    assert np.array_equal(got, np.repeat(logical, 8, axis=1))


def test_burst_rejects_short_rows(ctx):
    from ria_b200 import fec
    with pytest.raises(ValueError):
        fec.burst_deinterleave_batch(torch.zeros((1, 2, 2000), device="cuda"), ctx)
