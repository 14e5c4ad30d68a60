"""GPU parity of the BPF odometry (pf_odom_bpf_*, Odom_BPF_EstimationClass of the reference,
/root/reference/src/odomEstimationClass.cpp:649-1306) against the oracle restatement.  The reference feeds it beam / pillar /
facade clouds from its PCA feature extractor (out of scope); here the edge features are split into two line-type clouds and the
surf features play the facade cloud -- the arithmetic under test is the same."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _features(pfb, oracle, p, f):
    s = pfb.synth.scan(p, f)
    r = oracle.extract(s, num_lines=p.sensor_lines, order=1)
    e, u = s[r["edge_idx"]], s[r["surf_idx"]]
    return e[0::2], e[1::2], u      # beam, pillar, facade


@pytest.mark.parametrize("weight_type", [0.0, 12.0])
def test_bpf_sequence_matches_oracle(pfb, oracle, capi, weight_type):
    p = pfb.synth.config("cfg2")
    od = capi.OdometryBPF(0.4, 0, 0.4, 75, weight_type=weight_type, max_map_points=1 << 19)
    ref = oracle.OdomBPF(0.4, 0, 0.4, 75, weight_type)
    gp, rp = [], []
    for f in range(15):          # frames 11+ run as CUDA-graph replays (two kind pairs per update)
        b, pl, fa = _features(pfb, oracle, p, f)
        if f == 0:
            od.init_map(b, pl, fa)
            ref.init_map(b, pl, fa)
            continue
        gp.append(od.update(b, pl, fa))
        rp.append(ref.update(b, pl, fa))
    gp, rp = np.array(gp), np.array(rp)
    assert np.abs(gp[:, 4:] - rp[:, 4:]).max() < 2e-3
    assert np.abs(gp[:, :4] - rp[:, :4]).max() < 1e-4
    assert np.abs(gp[-1, 4:]).max() > 5.0                      # the vehicle moved ~14 m
    assert od.graph_captures >= 1
    np.testing.assert_allclose(od.iter_poses(), ref.iter_poses(), rtol=1e-3, atol=2e-4)
    rst = ref.stats()
    st = od.stats()
    assert st["n_edge_ds"] == rst["n_beam_ds"] and st["n_surf_ds"] == rst["n_facade_ds"] and st["passes"] == rst["passes"]
    assert abs(st["n_edge_res"] - rst["n_line_res"]) <= 0.01 * rst["n_line_res"] + 2
    assert abs(st["n_surf_res"] - rst["n_plane_res"]) <= 0.01 * rst["n_plane_res"] + 2
    for which in range(3):
        gm, rm = od.map_part(which), ref.get_map(which)
        assert abs(len(gm) - len(rm)) <= 0.01 * len(rm) + 2
    full = od.get_map()
    assert full.tobytes() == np.concatenate([od.map_part(0), od.map_part(1), od.map_part(2)]).tobytes()   # getMap order :683-689


def test_bpf_first_update_maps_are_identical(pfb, oracle, capi):
    p = pfb.synth.config("cfg2")
    od = capi.OdometryBPF(0.4, 0, 0.4, 75, max_map_points=1 << 19)
    ref = oracle.OdomBPF(0.4, 0, 0.4, 75)
    for f in range(2):
        b, pl, fa = _features(pfb, oracle, p, f)
        if f == 0:
            od.init_map(b, pl, fa); ref.init_map(b, pl, fa)
        else:
            od.update(b, pl, fa); ref.update(b, pl, fa)
    for which in range(3):
        gm, rm = od.map_part(which), ref.get_map(which)
        assert abs(len(gm) - len(rm)) <= 2
        if len(gm) == len(rm):
            assert (np.abs(gm["x"] - rm["x"]) < 1e-3).mean() > 0.999
            assert (gm["r"] == rm["r"]).mean() > 0.999


def test_bpf_handle_rejects_es_calls(capi):
    od = capi.OdometryBPF(max_map_points=65536, max_features=65536)
    z = np.zeros((10, 4), np.float32)
    with pytest.raises(capi.PfError):
        capi.Odometry.init_map(od, z, z)        # pf_odom_init_map on a 3-kind handle
