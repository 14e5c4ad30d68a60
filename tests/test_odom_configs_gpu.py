"""End-to-end parity on the other BASELINE.json configurations: residual weights (weightType 1 / 2 / 12),
the 32-ring sensor with PFilter 0/1/200 (configs[3]), F-LOAM mode 0/0/0 (configs[2]) and the 100-frame trajectory error
(north star: end-of-sequence ATE within 0.5 % of the reference's)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(pfb, oracle, capi, cfg, nframes, params, weight_type=0.0, stride=1):
    p = pfb.synth.config(cfg)
    k_new, theta_p, theta_max = params
    ex = capi.Extractor(num_lines=p.sensor_lines, max_points=131072)
    od = capi.Odometry(0.4, k_new, theta_p, theta_max, weight_type=weight_type, max_map_points=1 << 19)
    ref = oracle.Odom(0.4, k_new, theta_p, theta_max, weight_type)
    gp, rp = [], []
    for f in range(0, nframes * stride, stride):
        s = pfb.synth.scan(p, f)
        r = oracle.extract(s, num_lines=p.sensor_lines, order=1)
        pose = capi.frame_process(ex, od, s)
        if f == 0:
            ref.init_map(s[r["edge_idx"]], s[r["surf_idx"]])
            rpose = np.array([0, 0, 0, 1, 0, 0, 0.0])
        else:
            rpose = ref.update(s[r["edge_idx"]], s[r["surf_idx"]])
        gp.append(pose)
        rp.append(rpose)
    gt = np.array([pfb.synth.pose(p, f) for f in range(0, nframes * stride, stride)])
    return od, ref, np.array(gp), np.array(rp), gt


def _ate(poses, gt):
    rel = gt[:, 4:] - gt[0, 4:]
    return float(np.sqrt(((poses[:, 4:] - rel) ** 2).sum(1).mean()))


@pytest.mark.parametrize("weight_type", [1.0, 2.0, 12.0])
def test_residual_weights_match_oracle(pfb, oracle, capi, weight_type):
    od, ref, gp, rp, gt = _run(pfb, oracle, capi, "cfg2", 8, (0, 0.4, 75), weight_type)
    assert np.abs(gp[:, 4:] - rp[:, 4:]).max() < 2e-3
    assert np.abs(gp[:, :4] - rp[:, :4]).max() < 1e-4
    gi, ri = od.iter_poses(), ref.iter_poses()
    assert gi.shape == ri.shape
    np.testing.assert_allclose(gi, ri, rtol=1e-3, atol=2e-4)
    # the weights change the solution: the unweighted run differs measurably
    _, _, g0, _, _ = _run(pfb, oracle, capi, "cfg2", 8, (0, 0.4, 75), 0.0)
    assert np.abs(g0[:, 4:] - gp[:, 4:]).max() > 1e-5


def test_weight_type_validation(capi):
    with pytest.raises(capi.PfError) as e:
        capi.Odometry(0.4, 0, 0.4, 75, weight_type=3.0)
    assert e.value.status == -1


def test_vlp32_sequence_pfilter_0_1_200(pfb, oracle, capi):
    """configs[3] shape: 32-ring sensor, slow campus loop, PFilter 0/1/200 (README.md:43-44)."""
    od, ref, gp, rp, gt = _run(pfb, oracle, capi, "cfg4", 12, (0, 1.0, 200))
    assert np.abs(gp[:, 4:] - rp[:, 4:]).max() < 2e-3
    assert np.abs(gp[:, :4] - rp[:, :4]).max() < 1e-4
    st, rst = od.stats(), ref.stats()
    for k in ("n_edge_ds", "n_surf_ds", "passes"):
        assert st[k] == rst[k]
    assert abs(st["map_edge"] - rst["map_edge"]) <= 0.01 * rst["map_edge"] + 2
    assert abs(st["map_surf"] - rst["map_surf"]) <= 0.01 * rst["map_surf"] + 2


def test_trajectory_error_100_frames_within_half_percent(pfb, oracle, capi):
    """configs[1]: 100 frames, 0/0.4/75.  ATE against the generator's ground truth, GPU vs the CPU oracle."""
    od, ref, gp, rp, gt = _run(pfb, oracle, capi, "cfg2", 100, (0, 0.4, 75))
    a_gpu, a_ref = _ate(gp, gt), _ate(rp, gt)
    assert a_ref < 0.5                                    # the odometry tracks the 100 m trajectory
    assert abs(a_gpu - a_ref) <= 0.005 * a_ref, (a_gpu, a_ref)
    assert np.abs(gp[:, 4:] - rp[:, 4:]).max() < 5e-3


def test_floam_mode_100_frames(pfb, oracle, capi):
    """configs[2]: PFilter disabled (0, 0, 0) = the F-LOAM path; same bar."""
    od, ref, gp, rp, gt = _run(pfb, oracle, capi, "cfg2", 60, (0, 0.0, 0))
    a_gpu, a_ref = _ate(gp, gt), _ate(rp, gt)
    assert abs(a_gpu - a_ref) <= 0.005 * a_ref, (a_gpu, a_ref)
