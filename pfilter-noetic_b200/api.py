"""Host-side mirror of the reference's class interface for the hot path, over the C ABI (``capi``).

Same names, argument meaning, call order and error behaviour as the reference's classes, so that parity tests read like
a ROS-free harness of the reference:

  LaserProcessingClass   /root/reference/include/laserProcessingClass.h:30-42
  Odom_ES_EstimationClass (alias OdomEstimationClass)   /root/reference/include/odomEstimationClass.h:140-167
  Odom_BPF_EstimationClass  /root/reference/include/odomEstimationClass.h:169-205
  LaserMappingClass      /root/reference/include/laserMappingClass.h:32-58
  Lidar                  /root/reference/include/lidar.h:9-32

Clouds are numpy arrays: XYZI clouds are float32 [n, 4]; map / feature clouds with counters are ``capi.POINT_DTYPE``
records {x, y, z, r, g, b, a}.  Like the reference, the classes never raise on a bad frame: a failing call prints the
library's error text and carries on (src/odomEstimationClass.cpp:276,423,430 print and continue); the status of the
last call is kept in ``.status``.  There is no CPU path here: everything runs in libpfilter_b200.so on the GPU.
"""
import sys

import numpy as np

from . import capi


class Lidar:
    """lidar::Lidar (include/lidar.h:9-32): plain parameter block with the reference's setters."""

    def __init__(self):
        self.max_distance = 90.0
        self.min_distance = 3.0
        self.num_lines = 64
        self.scan_period = 0.1
        self.vertical_angle = 0.0
        self.vertical_angle_resolution = 0.0

    def setScanPeriod(self, v): self.scan_period = float(v)
    def setLines(self, v): self.num_lines = int(v)
    def setVerticalAngle(self, v): self.vertical_angle = float(v)
    def setVerticalResolution(self, v): self.vertical_angle_resolution = float(v)
    def setMaxDistance(self, v): self.max_distance = float(v)
    def setMinDistance(self, v): self.min_distance = float(v)


def _report(where, err):
    print(f"{where}: {err}", file=sys.stderr)


class LaserProcessingClass:
    """featureExtraction(pc_in, pc_out_edge, pc_out_surf) APPENDS to the caller's clouds (src/laserProcessingClass.cpp:10-96);
    here the two output clouds are Python lists of float32 [k, 4] chunks, or the call returns (edge, surf) arrays."""

    def __init__(self, device=0, max_points=131072):
        self._device, self._max_points = device, max_points
        self._ex = None
        self.status = 0

    def init(self, lidar_param):
        self.lidar_param = lidar_param
        self._ex = capi.Extractor(num_lines=lidar_param.num_lines, min_distance=lidar_param.min_distance,
                                  max_distance=lidar_param.max_distance, max_points=self._max_points, device=self._device)

    def featureExtraction(self, pc_in, pc_out_edge=None, pc_out_surf=None):
        try:
            edge, surf, _ = self._ex.run(pc_in, want_label=False)
            self.status = 0
        except capi.PfError as e:
            _report("featureExtraction", e)
            self.status = e.status
            edge = surf = np.zeros((0, 4), np.float32)
        if pc_out_edge is not None:
            pc_out_edge.append(edge)
        if pc_out_surf is not None:
            pc_out_surf.append(surf)
        return edge, surf

    @property
    def extractor(self):
        return self._ex


class Odom_ES_EstimationClass:
    """init / initMapWithPoints / updatePointsToMap / getMap with the public members the node reads after each call:
    ``odom`` (4x4 isometry, include/odomEstimationClass.h:57), laserCloudCornerMap / laserCloudSurfMap (:151-152; fetched
    from HBM on access)."""

    def __init__(self, device=0, max_map_points=0, max_features=0):
        self._device, self._mm, self._mf = device, max_map_points, max_features
        self._od = None
        self.odom = np.eye(4)
        self.pose7 = np.array([0, 0, 0, 1, 0, 0, 0.0])     # q_w_curr (x y z w), t_w_curr: `parameters` (:53-55)
        self.status = 0

    def init(self, lidar_param, map_resolution, k_new, theta_p, theta_max, weightType=0.0):
        self._od = capi.Odometry(map_resolution, k_new, theta_p, theta_max, weightType, self._mm, self._mf, self._device)

    def initMapWithPoints(self, edge_in, surf_in):
        try:
            self._od.init_map(edge_in, surf_in)
            self.status = 0
        except capi.PfError as e:
            _report("initMapWithPoints", e)
            self.status = e.status

    def updatePointsToMap(self, edge_in, surf_in):
        try:
            self._set_pose(self._od.update(edge_in, surf_in))
            self.status = 0
        except capi.PfError as e:
            _report("updatePointsToMap", e)
            self.status = e.status

    def getMap(self, laserCloudMap=None):
        """*laserCloudMap += surf map; += corner map (src/odomEstimationClass.cpp:210-215)."""
        m = self._od.get_map()
        if laserCloudMap is not None:
            laserCloudMap.append(m)
        return m

    @property
    def laserCloudCornerMap(self):
        return self._od.map_part(0)

    @property
    def laserCloudSurfMap(self):
        return self._od.map_part(1)

    def _set_pose(self, p):
        self.pose7 = np.asarray(p, float).copy()
        x, y, z, w = self.pose7[:4]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        self.odom = np.eye(4)
        self.odom[:3, :3] = R
        self.odom[:3, 3] = self.pose7[4:]

    @property
    def odometry(self):
        return self._od


OdomEstimationClass = Odom_ES_EstimationClass


class Odom_BPF_EstimationClass(Odom_ES_EstimationClass):
    """init / initMapWithPoints(beam, pillar, facade) / updatePointsToMap(beam, pillar, facade) / getMap
    (/root/reference/include/odomEstimationClass.h:169-205); maps: laserCloudBeamMap, laserCloudPillarMap, laserCloudFacadeMap."""

    def init(self, lidar_param, map_resolution, k_new, theta_p, theta_max, weightType=0.0):
        self._od = capi.OdometryBPF(map_resolution, k_new, theta_p, theta_max, weightType, self._mm, self._mf, self._device)

    def initMapWithPoints(self, beam_in, pillar_in, facade_in):
        try:
            self._od.init_map(beam_in, pillar_in, facade_in)
            self.status = 0
        except capi.PfError as e:
            _report("initMapWithPoints", e)
            self.status = e.status

    def updatePointsToMap(self, beam_in, pillar_in, facade_in):
        try:
            self._set_pose(self._od.update(beam_in, pillar_in, facade_in))
            self.status = 0
        except capi.PfError as e:
            _report("updatePointsToMap", e)
            self.status = e.status

    @property
    def laserCloudBeamMap(self):
        return self._od.map_part(0)

    @property
    def laserCloudPillarMap(self):
        return self._od.map_part(1)

    @property
    def laserCloudFacadeMap(self):
        return self._od.map_part(2)


class LaserMappingClass:
    """init(map_resolution) / updateCurrentPointsToMap(pc_in, pose_current) / getMap()
    (/root/reference/include/laserMappingClass.h:32-58).  pose_current is a 4x4 (or 3x4) isometry like the reference's
    Eigen::Isometry3d; the global map stays in HBM and getMap() copies it out on demand."""

    def __init__(self, device=0, max_map_points=0, max_points=0):
        self._device, self._mm, self._mp = device, max_map_points, max_points
        self._mp_handle = None
        self.status = 0

    def init(self, map_resolution):
        self._mp_handle = capi.Mapping(map_resolution, self._mm, self._mp, self._device)

    def updateCurrentPointsToMap(self, pc_in, pose_current):
        try:
            self._mp_handle.update(pc_in, np.asarray(pose_current, np.float64)[:3, :4])
            self.status = 0
        except capi.PfError as e:
            _report("updateCurrentPointsToMap", e)
            self.status = e.status

    def getMap(self):
        return self._mp_handle.get_map()

    @property
    def mapping(self):
        return self._mp_handle


def smoke_odometry(p, O):
    """Three frames of a synthetic sequence through the class interface on cuda:0, checked against the oracle
    (``O`` = the oracle module, passed in by __graft_entry__.smoke(); this module never imports it)."""
    from . import synth
    lid = Lidar()
    lid.setLines(p.sensor_lines)
    lp = LaserProcessingClass()
    lp.init(lid)
    od = Odom_ES_EstimationClass(max_map_points=262144)
    od.init(lid, 0.4, 0, 0.4, 75, 0.0)
    ref = O.Odom(0.4, 0, 0.4, 75)
    for f in range(3):
        s = synth.scan(p, f)
        e, u = lp.featureExtraction(s)
        r = O.extract(s, num_lines=p.sensor_lines, order=1)
        re, ru = s[r["edge_idx"]], s[r["surf_idx"]]
        assert np.array_equal(e, re) and np.array_equal(u, ru), "extraction differs from the oracle"
        if f == 0:
            od.initMapWithPoints(e, u)
            ref.init_map(re, ru)
        else:
            od.updatePointsToMap(e, u)
            rp = ref.update(re, ru)
            assert od.status == 0
            dq, dt = np.abs(od.pose7[:4] - rp[:4]).max(), np.abs(od.pose7[4:] - rp[4:]).max()
            assert dq < 1e-4 and dt < 2e-3, f"frame {f}: pose differs from the oracle (dq {dq:.2e}, dt {dt:.2e})"
    # global map (LaserMappingClass) and the BPF odometry class on the same data
    mp = LaserMappingClass(max_map_points=1 << 20, max_points=131072)
    mp.init(0.4)
    om = O.Mapping(0.4)
    for f in range(2):
        s = synth.scan(p, f)
        rt = capi.pose_to_rt(synth.pose(p, f))
        mp.updateCurrentPointsToMap(s, rt.reshape(3, 4))
        om.update(s, rt)
    gm, rm = mp.getMap(), om.get_map()
    assert mp.status == 0 and len(gm) == len(rm) and np.array_equal(np.sort(gm.view(np.uint32), axis=0), np.sort(rm.view(np.uint32), axis=0)), \
        "global map differs from the oracle"
    bpf = Odom_BPF_EstimationClass(max_map_points=262144)
    bpf.init(lid, 0.4, 0, 0.4, 75, 0.0)
    rb = O.OdomBPF(0.4, 0, 0.4, 75)
    for f in range(2):
        s = synth.scan(p, f)
        r = O.extract(s, num_lines=p.sensor_lines, order=1)
        e, u = s[r["edge_idx"]], s[r["surf_idx"]]
        if f == 0:
            bpf.initMapWithPoints(e[0::2], e[1::2], u)
            rb.init_map(e[0::2], e[1::2], u)
        else:
            bpf.updatePointsToMap(e[0::2], e[1::2], u)
            rp = rb.update(e[0::2], e[1::2], u)
            assert bpf.status == 0 and np.abs(bpf.pose7 - rp).max() < 2e-3, "BPF pose differs from the oracle"
    print(f"smoke: global map ok ({len(gm)} pts), BPF odometry ok")
    st = od.odometry.stats()
    print(f"smoke: odometry ok (3 frames, {od.odometry.launches} kernel launches, map {st['map_edge']}+{st['map_surf']} pts, "
          f"t = {np.round(od.pose7[4:], 3)})")
