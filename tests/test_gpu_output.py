"""Asynchronous output gather (kgpu_output_begin / kgpu_output_wait): the snapshot
delivered while the run continues must equal the blocking download at the same step, bit for bit,
and taking it must not disturb the run."""
import numpy as np
import pytest

from common import domain_stepper
from kestrel_b200 import capi
from kestrel_b200.host.synthetic import dambreak_runset, dambreak_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("morpho", [False, True])
def test_async_output_equals_blocking_download(morpho):
    import torch
    lib = capi.load_gpu()
    rs = dambreak_runset(4, 32, morpho=morpho)
    q4, b0v = dambreak_state(rs)
    a = domain_stepper(lib, rs, q4, b0v)     # takes asynchronous outputs on the way
    b = domain_stepper(lib, rs, q4, b0v)     # reference: blocking downloads
    NX, NY = rs.NX, rs.NY
    pinned = [torch.empty((4, NY, NX), dtype=torch.float64).pin_memory() for _ in range(2)]
    pinned_bt = [torch.empty((NY + 1, NX + 1), dtype=torch.float64).pin_memory() for _ in range(2)]
    snaps = []
    for interval in range(3):
        a.integrate_to(1e30, 7)
        b.integrate_to(1e30, 7)
        ref_q, ref_bt = b.download_domain(want_bt=True)
        buf, bbt = pinned[interval % 2].numpy(), pinned_bt[interval % 2].numpy()
        buf.fill(np.nan)
        a.output_begin(buf, bbt)             # returns at once; the next interval runs while it lands
        if interval == 1:                    # also legal: keep stepping before waiting
            a.integrate_to(1e30, 3)
            b.integrate_to(1e30, 3)
        a.output_wait()
        assert np.array_equal(buf, ref_q), f"interval {interval}"
        assert np.array_equal(bbt, ref_bt), f"interval {interval} (bed)"
        if morpho and interval == 2:
            assert np.max(np.abs(ref_bt)) > 0.0   # the bed did move: the comparison is not vacuous
        snaps.append(buf.copy())
    # the run that took outputs is still bit-identical to the one that did not
    assert np.array_equal(a.download_domain(), b.download_domain())
    assert not np.array_equal(snaps[0], snaps[2])
    a.output_wait()                          # idempotent with nothing in flight
    a.close(); b.close()
