"""Capacity / overflow conditions must surface as PF_ERR_CAPACITY from the fused frame path, not as a silently degraded pose
(the reference has no such limits: it prints and carries on, src/odomEstimationClass.cpp:276,423,430; ours are fixed at create
time and documented in include/pfilter_b200.h)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _feat(rng, n, half):
    a = np.zeros((n, 4), np.float32)
    a[:, :3] = (rng.random((n, 3), dtype=np.float32) - 0.5) * 2 * np.asarray(half, np.float32)
    return a


def test_ring_capacity_reaches_the_fused_frame_path(pfb, capi):
    """A ring with more points than max_ring_points is dropped whole by the extractor; pf_frame_process / pf_frame_wait must say so."""
    p = pfb.synth.config("cfg2")
    scans = [pfb.synth.scan(p, f) for f in range(3)]
    ex = capi.Extractor(num_lines=64, max_points=131072, max_ring_points=1024)     # the synthetic rings hold ~1790 points
    od = capi.Odometry(0.4, 0, 0.4, 75, max_map_points=1 << 19)
    with pytest.raises(capi.PfError) as e:
        capi.frame_process(ex, od, scans[0])
    assert e.value.status == -3 and "ring" in str(e.value)
    ex.close(); od.close()
    # queued form: the error comes out of pf_frame_wait
    ex = capi.Extractor(num_lines=64, max_points=131072, max_ring_points=1024)
    od = capi.Odometry(0.4, 0, 0.4, 75, max_map_points=1 << 19)
    ids = [capi.frame_submit(ex, od, s) for s in scans]
    with pytest.raises(capi.PfError) as e:
        capi.frame_wait(od, ids[1])
    assert e.value.status == -3
    # the stand-alone extraction call reports it as before
    with pytest.raises(capi.PfError):
        ex.run(scans[0])
    ex.close(); od.close()
    # the default capacity is fine
    ex = capi.Extractor(num_lines=64, max_points=131072)
    od = capi.Odometry(0.4, 0, 0.4, 75, max_map_points=1 << 19)
    for s in scans:
        capi.frame_process(ex, od, s)
    ex.close(); od.close()


def test_search_grid_overflow_is_reported(capi):
    """First-frame maps are not cropped (:217-222): an extent beyond 2^23 cells of 1 m cannot be searched."""
    rng = np.random.default_rng(5)
    od = capi.Odometry(0.4, 0, 0.4, 75, max_map_points=1 << 19)
    od.init_map(_feat(rng, 4000, (160, 160, 60)), _feat(rng, 20000, (160, 160, 60)))     # 321 x 321 x 121 cells > 2^23
    with pytest.raises(capi.PfError) as e:
        od.update(_feat(rng, 500, (20, 20, 2)), _feat(rng, 3000, (20, 20, 2)))
    assert e.value.status == -3 and "grid" in str(e.value)
    od.close()


def test_voxel_index_overflow_is_reported(capi):
    """pcl::VoxelGrid refuses index spaces beyond 2^31 ("leaf size too small"); the down-sampling kernels flag it."""
    rng = np.random.default_rng(6)
    od = capi.Odometry(0.4, 0, 0.4, 75, max_map_points=1 << 19)
    od.init_map(_feat(rng, 4000, (30, 30, 3)), _feat(rng, 20000, (30, 30, 3)))
    with pytest.raises(capi.PfError) as e:
        od.update(_feat(rng, 500, (400, 400, 400)), _feat(rng, 3000, (20, 20, 2)))       # 2000^3 voxels of 0.4 m
    assert e.value.status == -3
    od.close()


def test_outstanding_frame_limit_is_enforced(pfb, capi):
    """At most 32 submitted frames may await pf_frame_wait: the 33rd submit is refused instead of overwriting a result slot."""
    p = pfb.synth.config("cfg2")
    scan = pfb.synth.scan(p, 0)
    ex = capi.Extractor(num_lines=64, max_points=131072)
    od = capi.Odometry(0.4, 0, 0.4, 75, max_map_points=1 << 19)
    ids = [capi.frame_submit(ex, od, scan) for _ in range(32)]
    with pytest.raises(capi.PfError) as e:
        capi.frame_submit(ex, od, scan)
    assert e.value.status == -4
    capi.frame_wait(od, ids[0])
    ids.append(capi.frame_submit(ex, od, scan))      # one slot is free again
    for i in ids[1:]:
        capi.frame_wait(od, i)
    ex.close(); od.close()


def test_pose_history_range_copy(pfb, capi):
    p = pfb.synth.config("cfg2")
    ex = capi.Extractor(num_lines=64, max_points=131072)
    od = capi.Odometry(0.4, 0, 0.4, 75, max_map_points=1 << 19)
    poses = [capi.frame_process(ex, od, pfb.synth.scan(p, f)) for f in range(6)]
    import ctypes as C
    hist = np.zeros((5, 7))
    capi.check(capi.lib().pf_odom_get_pose_history(od.h, C.c_longlong(1), 5, hist.ctypes.data_as(C.c_void_p)))
    assert hist.tobytes() == np.array(poses[1:]).tobytes()
    ex.close(); od.close()


def test_non_finite_map_points_are_an_error_not_a_crash(capi, oracle):
    """A diverged pose turns appended points into inf / NaN; the search-grid build must flag them, not index with them."""
    rng = np.random.default_rng(8)
    xyz = (rng.random((20000, 3), dtype=np.float32) - 0.5) * np.array([20, 20, 4], np.float32)
    q = np.zeros((100, 4), np.float32)
    q[:, :3] = xyz[:100] + 0.05
    for bad in (np.inf, -np.inf, np.nan):
        m = xyz.copy()
        m[777, 1] = bad
        with pytest.raises(capi.PfError) as e:
            capi.knn5(capi.make_points(m), q)
        assert e.value.status == -3
    idx, d2 = capi.knn5(capi.make_points(xyz), q)          # the device is still healthy
    ci, cd = oracle.knn5(capi.make_points(xyz), q)
    assert np.array_equal(idx, ci) and (idx[:, 4] >= 0).all()


def test_tracking_loss_is_reported(pfb, capi):
    """The campus loop (cfg4) is where the reference algorithm itself loses its yaw estimate after ~540 frames; its constant-velocity
    prediction then doubles every frame until the pose is inf / NaN.  The GPU path must end that run with an error status."""
    from concurrent.futures import ThreadPoolExecutor
    p = pfb.synth.config("cfg4")
    pfb.synth.scan(p, 0)
    with ThreadPoolExecutor(8) as tp:
        scans = list(tp.map(lambda f: pfb.synth.scan(p, f), range(900)))
    ex = capi.Extractor(num_lines=32, max_points=57600)
    od = capi.Odometry(0.4, 0, 1.0, 200, max_map_points=1 << 20, max_features=57600)
    status = None
    for k, s in enumerate(scans):
        try:
            pose = capi.frame_process(ex, od, s)
        except capi.PfError as e:
            status = e.status
            break
        if not np.isfinite(pose).all():
            status = "nan pose returned"
            break
    # either the run survives (tracking kept) or it ends with a status -- never a CUDA fault (-2) and never a silent NaN
    assert status in (None, -3, -4), status
    ex.close(); od.close()
    ex = capi.Extractor(num_lines=32, max_points=57600)      # and the device is still usable
    ex.run(scans[0])
    ex.close()
