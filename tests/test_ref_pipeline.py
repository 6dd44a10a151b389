"""CPU: the oracle port of the whole path against the reference's OWN functions run end to end
(baseline/_ref, placed there by __graft_entry__.build() in the build container): per-pair match
counts and CRC32 of the (i, j) lists must be identical.  Skipped where baseline/_ref is absent."""

import numpy as np
import pytest
import torch

import oracle
from oracle import pipeline as opipe, ref_pipeline
from sslam_b200 import synth


@pytest.mark.skipif(not ref_pipeline.available(), reason="baseline/_ref not installed (run __graft_entry__.build())")
def test_oracle_pipeline_equals_reference_pipeline():
    from models.descriptor_refiner import DescriptorRefiner
    torch.manual_seed(0)
    refiner = DescriptorRefiner(384, 384, 128, 4)
    T, K = 5, 128
    sal, feat = synth.make_sequence(T, seq_id=3, height=96, width=128)
    sd = {k: v.detach().numpy() for k, v in refiner.state_dict().items()}
    ref_rec, _ = ref_pipeline.run_sequence(sal.numpy()[..., 0], feat.numpy(), sd, (384, 384, 128, 4), K, 2)
    w = oracle.RefinerWeights.from_state_dict(refiner.state_dict())
    port_rec, _ = opipe.run_sequence(sal.numpy()[..., 0], feat.numpy(), w, K, 1, 1)
    assert len(ref_rec) == T - 1
    assert [tuple(r) for r in ref_rec] == [tuple(r) for r in port_rec]
    assert all(r[0] > 0 for r in ref_rec)
