"""End-to-end: extraction + odometry + persistence filter on a synthetic sequence, GPU vs the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(pfb, oracle, capi, cfg, nframes, params, fused):
    p = pfb.synth.config(cfg)
    k_new, theta_p, theta_max = params
    ex = capi.Extractor(num_lines=p.sensor_lines, max_points=131072)
    od = capi.Odometry(0.4, k_new, theta_p, theta_max, max_map_points=262144)
    ref = oracle.Odom(0.4, k_new, theta_p, theta_max)
    gp, rp = [], []
    for f in range(nframes):
        s = pfb.synth.scan(p, f)
        r = oracle.extract(s, num_lines=p.sensor_lines, order=1)
        e, u = s[r["edge_idx"]], s[r["surf_idx"]]
        if fused:
            pose = capi.frame_process(ex, od, s)
        else:
            ge, gu, _ = ex.run(s, want_label=False)
            if f == 0:
                od.init_map(ge, gu)
                pose = od.pose()
            else:
                pose = od.update(ge, gu)
        if f == 0:
            ref.init_map(e, u)
            rpose = np.array([0, 0, 0, 1, 0, 0, 0.0])
        else:
            rpose = ref.update(e, u)
        gp.append(pose)
        rp.append(rpose)
    return p, od, ref, np.array(gp), np.array(rp)


@pytest.mark.parametrize("fused", [False, True])
def test_sequence_matches_oracle(pfb, oracle, capi, fused):
    n = 14
    p, od, ref, gp, rp = _run(pfb, oracle, capi, "cfg2", n, (0, 0.4, 75), fused)
    # per-frame poses within the stated tolerance (1e-4 relative on translation magnitudes of O(1..10) m)
    assert np.abs(gp[:, 4:] - rp[:, 4:]).max() < 2e-3
    assert np.abs(gp[:, :4] - rp[:, :4]).max() < 1e-4
    gt = np.array([pfb.synth.pose(p, f) for f in range(n)])
    assert np.abs((gp[:, 4:] - (gt[:, 4:] - gt[0, 4:]))).max() < 0.08      # and both track the ground truth
    st, rst = od.stats(), ref.stats()
    for k in ("n_edge_ds", "n_surf_ds", "passes"):
        assert st[k] == rst[k]
    assert abs(st["map_edge"] - rst["map_edge"]) <= 0.01 * rst["map_edge"]
    assert abs(st["map_surf"] - rst["map_surf"]) <= 0.01 * rst["map_surf"]
    assert od.launches > 50


def test_first_update_is_bit_identical_in_integer_outputs(pfb, oracle, capi):
    """Frame 1 starts from identical state on both sides: per-iteration poses agree to 1e-4 and the map keeps the
    same voxels (the only differences allowed are last-bit effects of sin/cos in the pose)."""
    p, od, ref, gp, rp = _run(pfb, oracle, capi, "cfg2", 2, (0, 0.4, 75), False)
    gi, ri = od.iter_poses(), ref.iter_poses()
    assert gi.shape == ri.shape == (11, 7)
    np.testing.assert_allclose(gi, ri, rtol=1e-4, atol=1e-6)
    for which in (0, 1):
        gm, rm = od.map_part(which), ref.get_map(which)
        assert abs(len(gm) - len(rm)) <= 2
        if len(gm) == len(rm):
            assert (np.abs(gm["x"] - rm["x"]) < 1e-3).mean() > 0.999
            assert (gm["r"] == rm["r"]).mean() > 0.999 and (gm["g"] == rm["g"]).mean() > 0.99


def test_getmap_order_and_pfilter_disabled(pfb, oracle, capi):
    p, od, ref, gp, rp = _run(pfb, oracle, capi, "cfg2", 4, (0, 0.0, 0), False)
    full = od.get_map()
    e, s = od.map_part(0), od.map_part(1)
    assert full.tobytes() == np.concatenate([s, e]).tobytes()      # getMap: surf first, then corner (:213-214)
    assert np.abs(gp[:, 4:] - rp[:, 4:]).max() < 2e-3


def test_update_before_init_is_an_error(capi):
    od = capi.Odometry(max_map_points=65536, max_features=65536)
    with pytest.raises(capi.PfError) as e:
        od.update(np.zeros((10, 4), np.float32), np.zeros((10, 4), np.float32))
    assert e.value.status == -4


def test_graph_replay_is_bit_identical_to_plain_launches(pfb, capi, monkeypatch):
    """From frame 11 on the update is replayed as a CUDA graph (csrc/odom.cu); the poses must not change by a single bit."""
    p = pfb.synth.config("cfg2")
    scans = [pfb.synth.scan(p, f) for f in range(24)]

    def run(graph):
        monkeypatch.setenv("PF_ODOM_GRAPH", graph)
        ex = capi.Extractor(num_lines=64, max_points=131072)
        od = capi.Odometry(0.4, 0, 0.4, 75, max_map_points=262144)
        poses = np.array([capi.frame_process(ex, od, s) for s in scans])
        maps = [od.map_part(0), od.map_part(1)]
        launches = od.launches
        ex.close(); od.close()
        return poses, maps, launches

    p1, m1, l1 = run("1")
    p0, m0, l0 = run("0")
    assert p1.tobytes() == p0.tobytes()
    assert m1[0].tobytes() == m0[0].tobytes() and m1[1].tobytes() == m0[1].tobytes()
    assert l1 == l0            # replayed kernels are counted like launched ones


def test_pipelined_submit_wait_equals_synchronous_calls(pfb, capi):
    """Queued frames run the down-sampling on its own stream (overlapping the previous frame's solve and map update) and, from
    frame 11 on, replay a graph without it; blocking calls keep it in line.  Poses and maps must not differ by a single bit."""
    p = pfb.synth.config("cfg2")
    scans = [pfb.synth.scan(p, f) for f in range(28)]
    ex, od = capi.Extractor(num_lines=64, max_points=131072), capi.Odometry(0.4, 0, 0.4, 75, max_map_points=262144)
    sync = np.array([capi.frame_process(ex, od, s) for s in scans])
    sync_maps = [od.map_part(0), od.map_part(1)]
    ex.close(); od.close()
    ex, od = capi.Extractor(num_lines=64, max_points=131072), capi.Odometry(0.4, 0, 0.4, 75, max_map_points=262144)
    ids = [capi.frame_submit(ex, od, s) for s in scans[:3]]          # three frames in flight
    out = []
    for k in range(3, len(scans)):
        ids.append(capi.frame_submit(ex, od, scans[k]))
        out.append(capi.frame_wait(od, ids[k - 3]))
    out += [capi.frame_wait(od, i) for i in ids[-3:]]
    assert ids == list(range(len(scans)))
    assert np.array(out).tobytes() == sync.tobytes()
    maps = [od.map_part(0), od.map_part(1)]
    assert maps[0].tobytes() == sync_maps[0].tobytes() and maps[1].tobytes() == sync_maps[1].tobytes()
    # mixing the two call styles on one handle (the graphs are re-captured for the other mode)
    more = [pfb.synth.scan(p, f) for f in range(28, 34)]
    a = [capi.frame_process(ex, od, s) for s in more[:3]]
    i3 = [capi.frame_submit(ex, od, s) for s in more[3:]]
    b = [capi.frame_wait(od, i) for i in i3]
    ex2, od2 = capi.Extractor(num_lines=64, max_points=131072), capi.Odometry(0.4, 0, 0.4, 75, max_map_points=262144)
    ref = np.array([capi.frame_process(ex2, od2, s) for s in scans + more])
    assert np.array(a + b).tobytes() == ref[28:].tobytes()
    ex2.close(); od2.close()
    with pytest.raises(capi.PfError):
        capi.frame_wait(od, 10_000)


def test_front_graph_recapture_new_extractor_and_several_handles(pfb, capi, monkeypatch):
    """pf_frame_submit replays extraction + down-sampling as a captured graph (csrc/odom.cu: front graph).  Scans that outgrow the
    size it was captured for, an extractor that is destroyed and replaced in mid-sequence, and a second sequence alive in the same
    process (programmatic launch edges then default to off) must not change a bit of the poses or the maps."""
    p = pfb.synth.config("cfg2")
    scans = [pfb.synth.scan(p, f) for f in range(26)]
    for k in range(12, 20):                       # shorter scans first, the full ones afterwards: the captured bound is outgrown
        scans[k] = np.ascontiguousarray(scans[k][: 60000 + 4000 * (k - 12)])

    def run(front, second_handle):
        monkeypatch.setenv("PF_FRAME_GRAPH", front)
        other = None
        if second_handle:
            other = (capi.Extractor(num_lines=64, max_points=131072), capi.Odometry(0.4, 0, 0.4, 75, max_map_points=262144))
        ex, od = capi.Extractor(num_lines=64, max_points=131072), capi.Odometry(0.4, 0, 0.4, 75, max_map_points=262144)
        out, ids = [], []
        for k, s in enumerate(scans):
            if k == 16:                           # the extractor is replaced: nothing captured with the old one may be replayed
                out += [capi.frame_wait(od, i) for i in ids]
                ids = []
                ex.close()
                ex = capi.Extractor(num_lines=64, max_points=131072)
            ids.append(capi.frame_submit(ex, od, s))
            if other is not None and k < 6:
                capi.frame_wait(other[1], capi.frame_submit(other[0], other[1], scans[k]))
            if len(ids) > 2:
                out.append(capi.frame_wait(od, ids.pop(0)))
        out += [capi.frame_wait(od, i) for i in ids]
        maps = [od.map_part(0), od.map_part(1)]
        ex.close(); od.close()
        if other is not None:
            other[0].close(); other[1].close()
        return np.array(out), maps

    a, ma = run("1", False)
    b, mb = run("0", False)
    c, mc = run("1", True)
    assert a.tobytes() == b.tobytes() == c.tobytes()
    for k in range(2):
        assert ma[k].tobytes() == mb[k].tobytes() == mc[k].tobytes()


def test_two_sequences_from_two_host_threads(pfb, capi):
    """configs[4] with fewer GPUs than sequences: several sequences share a GPU, each with its own handle pair, served by its own
    host thread (stream capture is thread-local, the library keeps no global mutable state besides the handle count).  Poses must
    equal those of the sequences run alone."""
    from concurrent.futures import ThreadPoolExecutor
    seqs = []
    for i in range(2):
        p = pfb.synth.config(f"cfg5.{i}")
        seqs.append([pfb.synth.scan(p, f) for f in range(24)])

    def run(scans):
        ex, od = capi.Extractor(num_lines=64, max_points=131072), capi.Odometry(0.4, 0, 0.4, 75, max_map_points=262144)
        ids, out = [], []
        for s in scans:
            ids.append(capi.frame_submit(ex, od, s))
            if len(ids) > 1:
                out.append(capi.frame_wait(od, ids.pop(0)))
        out += [capi.frame_wait(od, i) for i in ids]
        ex.close(); od.close()
        return np.array(out)

    alone = [run(s) for s in seqs]
    with ThreadPoolExecutor(2) as tp:
        together = list(tp.map(run, seqs))
    for a, b in zip(alone, together):
        assert a.tobytes() == b.tobytes()
