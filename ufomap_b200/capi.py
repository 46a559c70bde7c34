"""ctypes binding of libufomap_b200.so (include/ufomap_b200.h).

This is plumbing for tests/, bench.py and __graft_entry__: the product is the
shared library and its C ABI; the reference-facing host API is the C++ facade in
include/ufomap_b200/ufomap.hpp.  Loading fails loudly when the library has not
been built -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# UFOMAP_B200_LIB selects an experimental build of the same library (kernel variants)
LIB_PATH = os.environ.get("UFOMAP_B200_LIB") or os.path.join(HERE, "libufomap_b200.so")

OK, E_INVALID, E_CUDA, E_NOMEM, E_UNSUPPORTED = 0, 1, 2, 3, 4
XYZ_F64, XYZ_F32, XYZRGB_F64, XYZRGB_F32 = 0, 1, 2, 3


class Params(C.Structure):
    _fields_ = [
        ("resolution", C.c_double),
        ("depth_levels", C.c_uint32),
        ("automatic_pruning", C.c_int32),
        ("occupied_thres", C.c_double),
        ("free_thres", C.c_double),
        ("prob_hit", C.c_double),
        ("prob_miss", C.c_double),
        ("clamping_thres_min", C.c_double),
        ("clamping_thres_max", C.c_double),
        ("color", C.c_int32),
        ("device", C.c_int32),
        ("initial_blocks", C.c_uint64),
        ("initial_bricks", C.c_uint64),
    ]


class ScanStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "points", "rays", "visits", "touched_voxels", "hit_voxels", "touched_octets",
        "touched_blocks", "touched_d3", "touched_bricks", "upper_nodes", "blocks_in_map", "bricks_in_map",
        "device_bytes", "regrows", "launches", "result_bytes", "touched_lines")] + [(n, C.c_float) for n in (
            "ms_total", "ms_h2d", "ms_points", "ms_rays", "ms_scatter", "ms_update", "ms_propagate")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/ufomap_b200.h declares
SYMBOLS = [
    "ufo_b200_default_params", "ufo_b200_create", "ufo_b200_destroy", "ufo_b200_last_error",
    "ufo_b200_set_stream", "ufo_b200_insert_pointcloud", "ufo_b200_insert_device", "ufo_b200_wait",
    "ufo_b200_done", "ufo_b200_compute_ray", "ufo_b200_to_key", "ufo_b200_to_code",
    "ufo_b200_key_to_coord", "ufo_b200_key_to_code", "ufo_b200_code_to_key", "ufo_b200_query",
    "ufo_b200_export_leaves", "ufo_b200_set_sensor_model", "ufo_b200_sensor_model_logit",
    "ufo_b200_change_bbox", "ufo_b200_reset_change_bbox", "ufo_b200_last_scan_stats",
    "ufo_b200_set_profiling", "ufo_b200_clear", "ufo_b200_version", "ufo_b200_set_shard",
    "ufo_b200_insert_pointcloud_frame", "ufo_b200_transform_points", "ufo_b200_pose_from_rpy",
    "ufo_b200_insert_pointcloud2", "ufo_b200_write", "ufo_b200_write_file",
    "ufo_b200_write_data", "ufo_b200_set_value_volume", "ufo_b200_clear_resize",
    "ufo_b200_completed_scan_stats", "ufo_b200_set_sensor_model_field",
    "ufo_b200_route_inbox_bytes", "ufo_b200_route_setup", "ufo_b200_route_connect", "ufo_b200_route_mark",
    "ufo_b200_route_apply", "ufo_b200_ipc_export", "ufo_b200_ipc_open", "ufo_b200_ipc_close",
    "ufo_b200_write_compressed", "ufo_b200_cast_rays", "ufo_b200_export_nodes", "ufo_b200_read_data", "ufo_b200_enable_change_detection", "ufo_b200_reset_change_detection", "ufo_b200_changed_codes",
]

class Cloud2(C.Structure):
    """ufo_b200_cloud2: a sensor_msgs/PointCloud2 data buffer + field offsets."""
    _fields_ = [("data", C.c_void_p), ("n", C.c_size_t), ("point_step", C.c_uint32),
                ("off_x", C.c_uint32), ("off_y", C.c_uint32), ("off_z", C.c_uint32),
                ("off_r", C.c_int32), ("off_g", C.c_int32), ("off_b", C.c_int32),
                ("on_device", C.c_int32)]


_lib = None


def load():
    """Load the CUDA library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "ufomap_b200: %s is missing -- build it with `python -m ufomap_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, sz, dbl, u32, i32, u64p = C.c_void_p, C.c_size_t, C.c_double, C.c_uint32, C.c_int, C.c_void_p
    lib.ufo_b200_default_params.argtypes = [C.POINTER(Params)]
    lib.ufo_b200_default_params.restype = None
    lib.ufo_b200_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    lib.ufo_b200_destroy.argtypes = [vp]
    lib.ufo_b200_destroy.restype = None
    lib.ufo_b200_last_error.argtypes = [vp]
    lib.ufo_b200_last_error.restype = C.c_char_p
    lib.ufo_b200_set_stream.argtypes = [vp, vp]
    for f in (lib.ufo_b200_insert_pointcloud, lib.ufo_b200_insert_device):
        f.argtypes = [vp, vp, vp, sz, i32, dbl, u32, i32, u32, i32, i32]
    lib.ufo_b200_insert_pointcloud_frame.argtypes = [vp, vp, vp, sz, i32, vp, dbl, u32, i32, u32, i32,
                                                     i32]
    lib.ufo_b200_insert_pointcloud2.argtypes = [vp, vp, C.POINTER(Cloud2), vp, dbl, u32, i32, u32, i32, i32]
    lib.ufo_b200_write.argtypes = [vp, vp, u32, i32, vp, sz, C.POINTER(sz)]
    lib.ufo_b200_write_file.argtypes = [vp, C.c_char_p, vp, u32, i32]
    lib.ufo_b200_clear_resize.argtypes = [vp, dbl, u32]
    lib.ufo_b200_set_value_volume.argtypes = [vp, vp, dbl, u32]
    lib.ufo_b200_write_data.argtypes = [vp, vp, u32, vp, sz, C.POINTER(sz)]
    lib.ufo_b200_transform_points.argtypes = [vp, vp, sz, i32, vp]
    lib.ufo_b200_pose_from_rpy.argtypes = [dbl, dbl, dbl, dbl, dbl, dbl, vp]
    lib.ufo_b200_wait.argtypes = [vp]
    lib.ufo_b200_done.argtypes = [vp, C.POINTER(C.c_int)]
    lib.ufo_b200_compute_ray.argtypes = [vp, vp, vp, dbl, u32, u64p, sz, C.POINTER(sz)]
    lib.ufo_b200_to_key.argtypes = [vp, vp, u32, vp]
    lib.ufo_b200_to_code.argtypes = [vp, vp, u32, C.POINTER(C.c_uint64)]
    lib.ufo_b200_key_to_coord.argtypes = [vp, vp, u32, vp]
    lib.ufo_b200_key_to_code.argtypes = [vp]
    lib.ufo_b200_key_to_code.restype = C.c_uint64
    lib.ufo_b200_code_to_key.argtypes = [C.c_uint64, vp]
    lib.ufo_b200_code_to_key.restype = None
    lib.ufo_b200_query.argtypes = [vp, vp, vp, sz, vp, vp, vp]
    lib.ufo_b200_export_leaves.argtypes = [vp, vp, vp, vp, sz, C.POINTER(sz)]
    lib.ufo_b200_set_sensor_model.argtypes = [vp, vp]
    lib.ufo_b200_sensor_model_logit.argtypes = [vp, vp]
    lib.ufo_b200_change_bbox.argtypes = [vp, vp, vp]
    lib.ufo_b200_reset_change_bbox.argtypes = [vp]
    lib.ufo_b200_last_scan_stats.argtypes = [vp, C.POINTER(ScanStats)]
    lib.ufo_b200_completed_scan_stats.argtypes = [vp, C.POINTER(ScanStats)]
    lib.ufo_b200_set_sensor_model_field.argtypes = [vp, i32, dbl]
    lib.ufo_b200_set_profiling.argtypes = [vp, i32]
    lib.ufo_b200_write_compressed.argtypes = [vp, vp, u32, i32, i32, i32, vp, sz, C.POINTER(sz), C.POINTER(sz)]
    lib.ufo_b200_read_data.argtypes = [vp, vp, vp, sz, i32, sz]
    lib.ufo_b200_export_nodes.argtypes = [vp, u32, i32, i32, i32, vp, vp, vp, vp, sz, C.POINTER(sz)]
    lib.ufo_b200_cast_rays.argtypes = [vp, vp, vp, sz, i32, dbl, u32, vp, vp]
    lib.ufo_b200_enable_change_detection.argtypes = [vp, i32]
    lib.ufo_b200_reset_change_detection.argtypes = [vp]
    lib.ufo_b200_changed_codes.argtypes = [vp, u32, vp, sz, C.POINTER(sz)]
    lib.ufo_b200_route_inbox_bytes.argtypes = [u32, u32, u32, C.POINTER(sz)]
    lib.ufo_b200_route_setup.argtypes = [vp, u32, u32, u32, u32, C.POINTER(vp)]
    lib.ufo_b200_route_connect.argtypes = [vp, C.POINTER(vp)]
    lib.ufo_b200_route_mark.argtypes = [vp, vp, vp, sz, i32, dbl, i32, i32]
    lib.ufo_b200_route_apply.argtypes = [vp, u32, u32, i32]
    lib.ufo_b200_ipc_export.argtypes = [vp, vp]
    lib.ufo_b200_ipc_open.argtypes = [vp, C.POINTER(vp)]
    lib.ufo_b200_ipc_close.argtypes = [vp]
    lib.ufo_b200_clear.argtypes = [vp]
    lib.ufo_b200_set_shard.argtypes = [vp, u32, u32]
    lib.ufo_b200_version.restype = C.c_char_p
    _lib = lib
    return lib


class UfoError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("ufomap_b200 status %d: %s" % (status, msg))
        self.status = status


def pack_points(xyz, rgb=None, dtype=np.float64):
    """Pack points into one of the C-ABI layouts; returns (buffer, layout)."""
    xyz = np.asarray(xyz)
    n = len(xyz)
    if rgb is None:
        buf = np.ascontiguousarray(xyz, dtype=dtype).reshape(n, 3)
        return buf, (XYZ_F64 if dtype == np.float64 else XYZ_F32)
    rgb = np.asarray(rgb, dtype=np.uint8).reshape(n, 3)
    if dtype == np.float64:
        buf = np.zeros((n, 32), np.uint8)
        buf[:, :24] = np.ascontiguousarray(xyz, dtype=np.float64).reshape(n, 3).view(np.uint8).reshape(n, 24)
        buf[:, 24:27] = rgb
        return buf, XYZRGB_F64
    buf = np.zeros((n, 16), np.uint8)
    buf[:, :12] = np.ascontiguousarray(xyz, dtype=np.float32).reshape(n, 3).view(np.uint8).reshape(n, 12)
    buf[:, 12:15] = rgb
    return buf, XYZRGB_F32


class Map:
    """Thin object wrapper over the C ABI (one map = one device + one stream)."""

    def __init__(self, resolution, depth_levels=16, automatic_pruning=True, color=False,
                 device=-1, initial_blocks=0, initial_bricks=0, **model):
        self.lib = load()
        p = Params()
        self.lib.ufo_b200_default_params(C.byref(p))
        p.resolution = resolution
        p.depth_levels = depth_levels
        p.automatic_pruning = int(automatic_pruning)
        p.color = int(color)
        p.device = device
        p.initial_blocks = int(initial_blocks)
        p.initial_bricks = int(initial_bricks)
        for k, v in model.items():
            setattr(p, k, v)
        self.color = bool(color)
        self.resolution = resolution
        self.depth_levels = depth_levels
        self.h = C.c_void_p()
        rc = self.lib.ufo_b200_create(C.byref(p), C.byref(self.h))
        if rc == E_INVALID:
            raise ValueError("invalid map parameters (depth_levels has to be [2, 21])")
        if rc != OK:
            raise UfoError(rc, "ufo_b200_create failed (no usable CUDA device?)")

    def close(self):
        if getattr(self, "h", None):
            self.lib.ufo_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != OK:
            raise UfoError(rc, (self.lib.ufo_b200_last_error(self.h) or b"").decode())

    # -- integration --------------------------------------------------------
    def insert(self, origin, xyz, rgb=None, max_range=-1.0, depth=0, simple=False,
               early_stopping=0, discrete=False, async_=False, dtype=np.float64):
        buf, layout = pack_points(xyz, rgb, dtype)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        self._keep = (buf, o)
        self._check(self.lib.ufo_b200_insert_pointcloud(
            self.h, o.ctypes.data, buf.ctypes.data, len(buf), layout, float(max_range), int(depth),
            int(simple), int(early_stopping), int(discrete), int(async_)))

    def insert_frame(self, origin, xyz, frame_pose, rgb=None, max_range=-1.0, depth=0, simple=False,
                     discrete=False, async_=False, dtype=np.float64):
        """insertPointCloud(sensor_origin, cloud, frame_origin, ...): the cloud is moved into the
        map frame on the device.  frame_pose = (tx, ty, tz, qw, qx, qy, qz)."""
        buf, layout = pack_points(xyz, rgb, dtype)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        f = np.ascontiguousarray(frame_pose, dtype=np.float64)
        assert f.shape == (7,)
        self._keep = (buf, o, f)
        self._check(self.lib.ufo_b200_insert_pointcloud_frame(
            self.h, o.ctypes.data, buf.ctypes.data, len(buf), layout, f.ctypes.data,
            float(max_range), int(depth), int(simple), 0, int(discrete), int(async_)))

    def insert_pointcloud2(self, origin, data, point_step, off_xyz=(0, 4, 8), off_rgb=None,
                           frame_pose=None, max_range=-1.0, depth=0, simple=False, discrete=False,
                           async_=False):
        """Insert a raw sensor_msgs/PointCloud2 data buffer (numpy uint8 array); NaN points are
        skipped on the device like rosToUfo does."""
        buf = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        assert len(buf) % point_step == 0
        o = np.ascontiguousarray(origin, dtype=np.float64)
        f = None if frame_pose is None else np.ascontiguousarray(frame_pose, dtype=np.float64)
        r, g, b = (-1, -1, -1) if off_rgb is None else (off_rgb + 2, off_rgb + 1, off_rgb)
        c = Cloud2(buf.ctypes.data, len(buf) // point_step, point_step, off_xyz[0], off_xyz[1], off_xyz[2],
                   r, g, b, 0)
        self._keep = (buf, o, f, c)
        self._check(self.lib.ufo_b200_insert_pointcloud2(
            self.h, o.ctypes.data, C.byref(c), None if f is None else f.ctypes.data, float(max_range),
            int(depth), int(simple), 0, int(discrete), int(async_)))

    def insert_packed(self, origin, buf_ptr, n, layout, max_range=-1.0, depth=0, simple=False,
                      discrete=False, async_=False, on_device=False):
        """Insert from an already packed buffer (host pointer, or device pointer when
        on_device=True)."""
        o = np.ascontiguousarray(origin, dtype=np.float64)
        f = self.lib.ufo_b200_insert_device if on_device else self.lib.ufo_b200_insert_pointcloud
        self._check(f(self.h, o.ctypes.data, buf_ptr, int(n), int(layout), float(max_range),
                      int(depth), int(simple), 0, int(discrete), int(async_)))

    def wait(self):
        self._check(self.lib.ufo_b200_wait(self.h))

    def done(self):
        d = C.c_int()
        self._check(self.lib.ufo_b200_done(self.h, C.byref(d)))
        return bool(d.value)

    def set_stream(self, stream_ptr):
        self._check(self.lib.ufo_b200_set_stream(self.h, stream_ptr))

    def set_profiling(self, enable=True):
        self._check(self.lib.ufo_b200_set_profiling(self.h, int(enable)))

    def stats(self):
        st = ScanStats()
        self._check(self.lib.ufo_b200_last_scan_stats(self.h, C.byref(st)))
        return st.as_dict()

    def completed_stats(self):
        """Statistics of the last COMPLETED scan (no wait for the scan in flight)."""
        st = ScanStats()
        self._check(self.lib.ufo_b200_completed_scan_stats(self.h, C.byref(st)))
        return st.as_dict()

    def set_sensor_model_field(self, index, probability):
        """One parameter, like the reference's setProbHit/... (the other stored log-odds stay)."""
        self._check(self.lib.ufo_b200_set_sensor_model_field(self.h, int(index), float(probability)))

    def clear(self):
        self._check(self.lib.ufo_b200_clear(self.h))

    def set_shard(self, rank, world):
        """Keep only the bricks this rank owns (spatial sharding over several GPUs)."""
        self._check(self.lib.ufo_b200_set_shard(self.h, int(rank), int(world)))

    # -- routed multi-GPU mode ----------------------------------------------
    def route_setup(self, rank, world, cap_bricks=1 << 16, cap_hits=1 << 18):
        """Allocate this rank's inbox; returns its device pointer (int)."""
        ptr = C.c_void_p()
        self._check(self.lib.ufo_b200_route_setup(self.h, int(rank), int(world), int(cap_bricks), int(cap_hits),
                                                  C.byref(ptr)))
        self.route_world = world
        return ptr.value

    def route_connect(self, peer_ptrs):
        arr = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(p) for p in peer_ptrs])
        self._check(self.lib.ufo_b200_route_connect(self.h, arr))

    def route_mark(self, origin, buf_ptr, n, layout, max_range=-1.0, on_device=False, self_too=False):
        o = np.ascontiguousarray(origin, dtype=np.float64)
        self._check(self.lib.ufo_b200_route_mark(self.h, o.ctypes.data, buf_ptr, int(n), int(layout),
                                                 float(max_range), int(on_device), int(self_too)))

    def route_apply(self, first=0, count=None, last=True):
        count = self.route_world if count is None else count
        self._check(self.lib.ufo_b200_route_apply(self.h, int(first), int(count), int(last)))

    # -- file / wire format -------------------------------------------------
    @staticmethod
    def _box6(box):
        """(min xyz, max xyz) -> the AABB's centre / half size, like its (min, max) constructor."""
        if box is None:
            return None
        mn, mx = np.asarray(box[0], np.float64), np.asarray(box[1], np.float64)
        half = (mx - mn) / 2.0
        return np.ascontiguousarray(np.concatenate([mn + half, half]), np.float64)

    def write(self, expanded=False, box=None, min_depth=0):
        """The map as a UFOMap file image (bytes), Octree::write with compress=False."""
        b = self._box6(box)
        bp = None if b is None else b.ctypes.data
        n = C.c_size_t()
        self._check(self.lib.ufo_b200_write(self.h, bp, int(min_depth), int(expanded), None, 0, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        self._check(self.lib.ufo_b200_write(self.h, bp, int(min_depth), int(expanded), buf.ctypes.data, n.value,
                                            C.byref(n)))
        assert n.value == len(buf)
        return buf.tobytes()

    def write_compressed(self, box=None, min_depth=0, data_only=False, acceleration=1, level=0):
        """Octree::write / writeData with compress=True: (bytes, uncompressed_data_size)."""
        b = self._box6(box)
        bp = None if b is None else b.ctypes.data
        n, u = C.c_size_t(), C.c_size_t()
        self._check(self.lib.ufo_b200_write_compressed(self.h, bp, int(min_depth), int(data_only), int(acceleration),
                                                       int(level), None, 0, C.byref(n), C.byref(u)))
        buf = np.empty(max(n.value, 1), np.uint8)
        self._check(self.lib.ufo_b200_write_compressed(self.h, bp, int(min_depth), int(data_only), int(acceleration),
                                                       int(level), buf.ctypes.data, n.value, C.byref(n), C.byref(u)))
        return buf[:n.value].tobytes(), int(u.value)

    def read_data(self, data, box=None, compressed=False, uncompressed_size=0):
        """Octree::readData: merge a node stream (bytes) into the map; box = (min xyz, max xyz)."""
        b = self._box6(box)
        buf = np.frombuffer(data, np.uint8)
        self._check(self.lib.ufo_b200_read_data(self.h, None if b is None else b.ctypes.data,
                                                buf.ctypes.data if len(buf) else None, len(buf), int(compressed),
                                                int(uncompressed_size)))

    def write_file(self, filename, expanded=False, box=None, min_depth=0):
        b = self._box6(box)
        self._check(self.lib.ufo_b200_write_file(self.h, os.fsencode(filename),
                                                 None if b is None else b.ctypes.data, int(min_depth),
                                                 int(expanded)))

    def clear_resize(self, resolution, depth_levels):
        """Octree::clear(resolution, depth_levels)."""
        self._check(self.lib.ufo_b200_clear_resize(self.h, float(resolution), int(depth_levels)))
        self.resolution, self.depth_levels = resolution, depth_levels

    def set_value_volume(self, box, occupancy, min_depth=0):
        """setValueVolume(AABB(min, max), occupancy probability, min_depth); box = (min xyz, max xyz)."""
        b = self._box6(box)
        self._check(self.lib.ufo_b200_set_value_volume(self.h, b.ctypes.data, float(occupancy), int(min_depth)))

    def write_data(self, box=None, min_depth=0):
        """Octree::writeData(stream, AABB(min, max) or whole map, False, min_depth): node stream.
        box = (min xyz, max xyz); converted to the AABB's centre / half size like its constructor."""
        b = self._box6(box)
        bp = None if b is None else b.ctypes.data
        n = C.c_size_t()
        self._check(self.lib.ufo_b200_write_data(self.h, bp, int(min_depth), None, 0, C.byref(n)))
        buf = np.empty(max(n.value, 1), np.uint8)
        self._check(self.lib.ufo_b200_write_data(self.h, bp, int(min_depth), buf.ctypes.data, n.value,
                                                 C.byref(n)))
        return buf[:n.value].tobytes()

    # -- state ----------------------------------------------------------------
    def value_field(self, sort=True):
        """(codes sorted, occ f32, rgb u8[n,3]): every voxel with a non-default payload.
        sort=False returns them in the device's emission order."""
        n = C.c_size_t()
        self._check(self.lib.ufo_b200_export_leaves(self.h, None, None, None, 0, C.byref(n)))
        cnt = n.value
        codes = np.empty(cnt, np.uint64)
        occ = np.empty(cnt, np.float32)
        rgb = np.zeros((cnt, 3), np.uint8)
        if cnt:
            self._check(self.lib.ufo_b200_export_leaves(
                self.h, codes.ctypes.data, occ.ctypes.data, rgb.ctypes.data if self.color else None,
                cnt, C.byref(n)))
            assert n.value == cnt
        if not sort:
            return codes, occ, rgb
        order = np.argsort(codes, kind="stable")
        return codes[order], occ[order], rgb[order]

    def value_field_count(self):
        """Number of voxels value_field() would return."""
        n = C.c_size_t()
        self._check(self.lib.ufo_b200_export_leaves(self.h, None, None, None, 0, C.byref(n)))
        return int(n.value)

    def query(self, codes, depths):
        codes = np.ascontiguousarray(codes, dtype=np.uint64)
        depths = np.ascontiguousarray(np.broadcast_to(depths, codes.shape), dtype=np.uint32)
        n = len(codes)
        occ = np.empty(n, np.float32)
        flags = np.empty(n, np.uint8)
        rgb = np.zeros((n, 3), np.uint8)
        self._check(self.lib.ufo_b200_query(self.h, codes.ctypes.data, depths.ctypes.data, n,
                                            occ.ctypes.data, flags.ctypes.data,
                                            rgb.ctypes.data if self.color else None))
        return occ, flags, rgb

    def export_nodes(self, depth=0, occupied=True, free=True, unknown=False, box=None):
        """Filtered leaf iteration as a batch: (codes sorted, occ, rgb) of the depth-`depth` nodes
        whose state passes the filter and whose cube intersects box = (min xyz, max xyz)."""
        b = self._box6(box)
        bp = None if b is None else b.ctypes.data
        n = C.c_size_t()
        self._check(self.lib.ufo_b200_export_nodes(self.h, int(depth), int(occupied), int(free), int(unknown), bp,
                                                   None, None, None, 0, C.byref(n)))
        cnt = n.value
        codes = np.empty(cnt, np.uint64)
        occ = np.empty(cnt, np.float32)
        rgb = np.zeros((cnt, 3), np.uint8)
        if cnt:
            self._check(self.lib.ufo_b200_export_nodes(self.h, int(depth), int(occupied), int(free), int(unknown), bp,
                                                       codes.ctypes.data, occ.ctypes.data,
                                                       rgb.ctypes.data if self.color else None, cnt, C.byref(n)))
        order = np.argsort(codes, kind="stable")
        return codes[order], occ[order], rgb[order]

    def cast_rays(self, origins, directions, ignore_unknown=False, max_range=-1.0, depth=0):
        """Batched castRay: (hit bool[n], code u64[n])."""
        o = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
        assert o.shape == d.shape
        codes = np.zeros(len(o), np.uint64)
        hit = np.zeros(len(o), np.uint8)
        self._check(self.lib.ufo_b200_cast_rays(self.h, o.ctypes.data, d.ctypes.data, len(o), int(ignore_unknown),
                                                float(max_range), int(depth), codes.ctypes.data, hit.ctypes.data))
        return hit.astype(bool), codes

    # -- geometry -----------------------------------------------------------
    def compute_ray(self, origin, end, max_range=-1.0, depth=0):
        o = np.ascontiguousarray(origin, dtype=np.float64)
        e = np.ascontiguousarray(end, dtype=np.float64)
        n = C.c_size_t()
        self._check(self.lib.ufo_b200_compute_ray(self.h, o.ctypes.data, e.ctypes.data,
                                                  float(max_range), int(depth), None, 0, C.byref(n)))
        out = np.empty(n.value, np.uint64)
        if n.value:
            self._check(self.lib.ufo_b200_compute_ray(self.h, o.ctypes.data, e.ctypes.data,
                                                      float(max_range), int(depth), out.ctypes.data,
                                                      n.value, C.byref(n)))
        return out

    def to_key(self, xyz, depth=0):
        p = np.ascontiguousarray(xyz, dtype=np.float64)
        k = np.empty(3, np.uint32)
        self._check(self.lib.ufo_b200_to_key(self.h, p.ctypes.data, int(depth), k.ctypes.data))
        return k

    def to_code(self, xyz, depth=0):
        p = np.ascontiguousarray(xyz, dtype=np.float64)
        c = C.c_uint64()
        self._check(self.lib.ufo_b200_to_code(self.h, p.ctypes.data, int(depth), C.byref(c)))
        return int(c.value)

    def key_to_code(self, key, depth=0):
        k = np.ascontiguousarray(key, dtype=np.uint32)
        return int(self.lib.ufo_b200_key_to_code(k.ctypes.data))

    def code_to_key(self, code, depth=0):
        k = np.empty(3, np.uint32)
        self.lib.ufo_b200_code_to_key(int(code), k.ctypes.data)
        return k

    def key_to_coord(self, key, depth=0):
        k = np.ascontiguousarray(key, dtype=np.uint32)
        p = np.empty(3, np.float64)
        self._check(self.lib.ufo_b200_key_to_coord(self.h, k.ctypes.data, int(depth), p.ctypes.data))
        return p

    def change_bbox(self):
        mn, mx = np.empty(3), np.empty(3)
        self._check(self.lib.ufo_b200_change_bbox(self.h, mn.ctypes.data, mx.ctypes.data))
        return mn, mx

    def enable_change_detection(self, enable=True):
        self._check(self.lib.ufo_b200_enable_change_detection(self.h, int(enable)))

    def reset_change_detection(self):
        self._check(self.lib.ufo_b200_reset_change_detection(self.h))

    def changed_codes(self, depth=0):
        """Sorted codes (at `depth`) of the nodes whose value changed since the last reset."""
        n = C.c_size_t()
        self._check(self.lib.ufo_b200_changed_codes(self.h, int(depth), None, 0, C.byref(n)))
        out = np.empty(n.value, np.uint64)
        if n.value:
            self._check(self.lib.ufo_b200_changed_codes(self.h, int(depth), out.ctypes.data, n.value, C.byref(n)))
        return np.sort(out[:n.value])

    def reset_change_bbox(self):
        self._check(self.lib.ufo_b200_reset_change_bbox(self.h))

    def sensor_model(self):
        out = np.empty(6)
        self._check(self.lib.ufo_b200_sensor_model_logit(self.h, out.ctypes.data))
        return out

    def set_sensor_model(self, occupied_thres=0.5, free_thres=0.5, prob_hit=0.7, prob_miss=0.4,
                         clamping_thres_min=0.1192, clamping_thres_max=0.971):
        p = np.array([occupied_thres, free_thres, prob_hit, prob_miss, clamping_thres_min,
                      clamping_thres_max], dtype=np.float64)
        self._check(self.lib.ufo_b200_set_sensor_model(self.h, p.ctypes.data))


def transform_points(frame_pose, xyz, dtype=np.float64):
    """Pose6::transform on the host (bit-identical to the device path)."""
    lib = load()
    buf, layout = pack_points(xyz, None, dtype)
    f = np.ascontiguousarray(frame_pose, dtype=np.float64)
    out = np.empty((len(buf), 3), np.float64)
    rc = lib.ufo_b200_transform_points(f.ctypes.data, buf.ctypes.data, len(buf), layout,
                                       out.ctypes.data)
    if rc != 0:
        raise RuntimeError("ufo_b200_transform_points failed: %d" % rc)
    return out


def pose_from_rpy(x, y, z, roll, pitch, yaw):
    """Pose6(x, y, z, roll, pitch, yaw) as (tx, ty, tz, qw, qx, qy, qz)."""
    lib = load()
    out = np.empty(7, np.float64)
    lib.ufo_b200_pose_from_rpy(float(x), float(y), float(z), float(roll), float(pitch), float(yaw),
                               out.ctypes.data)
    return out


def ipc_export(dev_ptr):
    """cudaIpcMemHandle (64 bytes) of a device allocation of this process."""
    lib = load()
    h = (C.c_ubyte * 64)()
    rc = lib.ufo_b200_ipc_export(C.c_void_p(dev_ptr), h)
    if rc != 0:
        raise UfoError(rc, "cudaIpcGetMemHandle failed")
    return bytes(h)


def ipc_open(handle):
    """Map another process's allocation; returns the device pointer (int) valid in this process."""
    lib = load()
    buf = (C.c_ubyte * 64).from_buffer_copy(handle)
    p = C.c_void_p()
    rc = lib.ufo_b200_ipc_open(buf, C.byref(p))
    if rc != 0:
        raise UfoError(rc, "cudaIpcOpenMemHandle failed")
    return p.value


def ipc_close(dev_ptr):
    """Unmap a pointer obtained from ipc_open."""
    lib = load()
    rc = lib.ufo_b200_ipc_close(C.c_void_p(dev_ptr))
    if rc != 0:
        raise UfoError(rc, "cudaIpcCloseMemHandle failed")
