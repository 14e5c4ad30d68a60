"""GPU parity: pf_extract_* (CUDA) vs the extraction oracle (restatement, itself pinned to the real reference)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(scan, out, ref):
    edge, surf, label = out
    assert np.array_equal(label, ref["label"])
    assert np.array_equal(edge, scan[ref["edge_idx"]])          # bit-exact points, same emission order
    assert np.array_equal(surf, scan[ref["surf_idx"]])


def test_extract_single_matches_oracle(capi, oracle, cfg2_scans):
    _, scans = cfg2_scans
    ex = capi.Extractor(num_lines=64, max_points=131072)
    for s in scans:
        _check(s, ex.run(s), oracle.extract(s, order=1))
    assert ex.launches >= 2 * len(scans)


def test_extract_sets_match_real_reference(capi, oracle, cfg2_scans):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    _, scans = cfg2_scans
    ex = capi.Extractor(num_lines=64, max_points=131072)
    s = scans[1]
    _, _, label = ex.run(s)
    e, f = oracle.ref_extract(s)
    ref_label = np.zeros(len(s), np.uint8)
    ref_label[f] = 2
    ref_label[e] = 1
    assert np.array_equal(label, ref_label)


def test_extract_batch_and_ragged(capi, oracle, pfb):
    p = pfb.synth.config("cfg2")
    scans = [pfb.synth.scan(p, f) for f in (0, 7, 33)]
    scans.append(scans[0][:50000])               # truncated scan: last ring cut short
    scans.append(scans[1][:100])                 # fewer than 131 points in every ring -> nothing selected
    scans.append(np.zeros((0, 4), np.float32))   # empty scan
    ex = capi.Extractor(num_lines=64, max_points=115200, max_batch=8)
    outs = ex.run_batch(scans)
    for s, o in zip(scans, outs):
        _check(s, o, oracle.extract(s, order=1))
    assert len(outs[-1][0]) == 0 and len(outs[-2][1]) == 0


def test_extract_interleaved_point_order(capi, oracle, cfg2_scans):
    """Input that is NOT ring-major (azimuth-major interleave): the stable per-ring gather must still hold."""
    _, scans = cfg2_scans
    s = scans[0]
    ring = oracle.extract(s)["ring"]
    # stable reorder by position-in-ring: point k of every ring, then point k+1 ...
    pos = np.zeros(len(s), np.int64)
    for r in range(64):
        idx = np.nonzero(ring == r)[0]
        pos[idx] = np.arange(len(idx))
    perm = np.lexsort((ring, pos))
    t = np.ascontiguousarray(s[perm])
    ex = capi.Extractor(num_lines=64, max_points=131072)
    _check(t, ex.run(t), oracle.extract(t, order=1))


def test_extract_32_and_16_lines(capi, oracle, pfb):
    for lines in (32, 16):
        p = pfb.synth.params(sensor_lines=lines, seed=77)
        s = pfb.synth.scan(p, 3)
        ex = capi.Extractor(num_lines=lines, max_points=65536)
        _check(s, ex.run(s), oracle.extract(s, num_lines=lines, order=1))


def test_extract_ring_capacity_error(capi, cfg2_scans):
    _, scans = cfg2_scans
    ex = capi.Extractor(num_lines=64, max_points=131072, max_ring_points=1024)
    with pytest.raises(capi.PfError) as e:
        ex.run(scans[0])
    assert e.value.status == -3
