"""COLMAP binary model I/O (cameras.bin / images.bin / points3D.bin) for the HP2 boundary.

The pipeline's output format stays what the reference writes and reads through
`sfm/colmap_utils/read_write_model.py` (`read_model` :419-445, `write_model` :447-456) — the
COLMAP sparse-model layout.  This module reads such a model straight into the
`particlesfm_b200.ba.Reconstruction` the bundle adjuster consumes and writes it back,
with bulk `numpy.frombuffer` decoding of the per-image observation lists and per-point
tracks instead of one `struct.unpack` per element.

Layout (little endian), as documented by COLMAP:
  cameras.bin   u64 n | per camera: i32 id, i32 model, u64 width, u64 height, f64 params[k(model)]
  images.bin    u64 n | per image:  i32 id, f64 q[4], f64 t[3], i32 camera_id, name '\\0',
                                    u64 m, m x (f64 x, f64 y, i64 point3D_id)
  points3D.bin  u64 n | per point:  i64 id, f64 xyz[3], u8 rgb[3], f64 error,
                                    u64 l, l x (i32 image_id, i32 point2D_idx)

tests/test_colmap_io.py checks both directions against files written / parsed by the
reference's own module (tests/golden/colmap_model/, tests/golden/make_colmap_golden.py):
reading them gives the generating values, writing the same model reproduces them byte for byte.
"""
import os
import struct

import numpy as np

from .ba import Camera, Image, Point3D, Reconstruction

# number of parameters per COLMAP camera model id
NUM_PARAMS = {0: 3, 1: 4, 2: 4, 3: 5, 4: 8, 5: 8, 6: 12, 7: 5, 8: 4, 9: 5, 10: 12}
MODEL_NAMES = {0: "SIMPLE_PINHOLE", 1: "PINHOLE", 2: "SIMPLE_RADIAL", 3: "RADIAL", 4: "OPENCV", 5: "OPENCV_FISHEYE",
               6: "FULL_OPENCV", 7: "FOV", 8: "SIMPLE_RADIAL_FISHEYE", 9: "RADIAL_FISHEYE", 10: "THIN_PRISM_FISHEYE"}

_P2D = np.dtype([("x", "<f8"), ("y", "<f8"), ("id", "<i8")])
_TRK = np.dtype([("image_id", "<i4"), ("point2D_idx", "<i4")])


def read_cameras_bin(path):
    buf = open(path, "rb").read()
    (n,), o = struct.unpack_from("<Q", buf, 0), 8
    cams = {}
    for _ in range(n):
        cid, model, w, h = struct.unpack_from("<iiQQ", buf, o)
        o += 24
        if model not in NUM_PARAMS:
            raise ValueError(f"cameras.bin: unknown camera model id {model}")
        k = NUM_PARAMS[model]
        cams[cid] = Camera(cid, model, int(w), int(h), np.frombuffer(buf, "<f8", k, o).copy())
        o += 8 * k
    return cams


def read_images_bin(path):
    buf = open(path, "rb").read()
    (n,), o = struct.unpack_from("<Q", buf, 0), 8
    images = {}
    for _ in range(n):
        iid = struct.unpack_from("<i", buf, o)[0]
        qt = np.frombuffer(buf, "<f8", 7, o + 4)
        cam = struct.unpack_from("<i", buf, o + 60)[0]
        o += 64
        e = buf.index(b"\x00", o)
        name = buf[o:e].decode("utf-8")
        o = e + 1
        (m,) = struct.unpack_from("<Q", buf, o)
        o += 8
        p = np.frombuffer(buf, _P2D, m, o)
        o += 24 * m
        images[iid] = Image(iid, qt[:4].copy(), qt[4:].copy(), cam, name,
                            np.stack([p["x"], p["y"]], axis=1) if m else np.zeros((0, 2)), p["id"].astype(np.int64))
    return images


def read_points3D_bin(path):
    buf = open(path, "rb").read()
    (n,), o = struct.unpack_from("<Q", buf, 0), 8
    pts = {}
    for _ in range(n):
        pid = struct.unpack_from("<q", buf, o)[0]
        xyz = np.frombuffer(buf, "<f8", 3, o + 8).copy()
        rgb = np.frombuffer(buf, np.uint8, 3, o + 32).copy()
        err, l = struct.unpack_from("<dQ", buf, o + 35)
        o += 51
        t = np.frombuffer(buf, _TRK, l, o)
        o += 8 * l
        pts[pid] = Point3D(pid, xyz, rgb, err, t["image_id"].astype(np.int32), t["point2D_idx"].astype(np.int32))
    return pts


def read_model(path):
    """cameras.bin + images.bin + points3D.bin in `path` -> Reconstruction."""
    return Reconstruction(read_cameras_bin(os.path.join(path, "cameras.bin")),
                          read_images_bin(os.path.join(path, "images.bin")),
                          read_points3D_bin(os.path.join(path, "points3D.bin")))


def write_cameras_bin(cameras, path):
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(cameras)))
        for c in cameras.values():
            f.write(struct.pack("<iiQQ", c.camera_id, c.model_id, c.width, c.height))
            f.write(np.asarray(c.params, "<f8").tobytes())


def write_images_bin(images, path):
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(images)))
        for im in images.values():
            f.write(struct.pack("<i", im.image_id))
            f.write(np.asarray(im.qvec, "<f8").tobytes())
            f.write(np.asarray(im.tvec, "<f8").tobytes())
            f.write(struct.pack("<i", im.camera_id))
            f.write(im.name.encode("utf-8") + b"\x00")
            m = 0 if im.xys is None else len(im.xys)
            f.write(struct.pack("<Q", m))
            if m:
                p = np.empty(m, _P2D)
                p["x"], p["y"], p["id"] = im.xys[:, 0], im.xys[:, 1], im.point3D_ids
                f.write(p.tobytes())


def write_points3D_bin(points3D, path):
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(points3D)))
        for p in points3D.values():
            f.write(struct.pack("<q", p.point3D_id))
            f.write(np.asarray(p.xyz, "<f8").tobytes())
            f.write(np.asarray(p.rgb if p.rgb is not None else (0, 0, 0), np.uint8).tobytes())
            l = 0 if p.image_ids is None else len(p.image_ids)
            f.write(struct.pack("<dQ", p.error, l))
            if l:
                t = np.empty(l, _TRK)
                t["image_id"], t["point2D_idx"] = p.image_ids, p.point2D_idxs
                f.write(t.tobytes())


def write_model(reconstruction, path):
    os.makedirs(path, exist_ok=True)
    write_cameras_bin(reconstruction.cameras, os.path.join(path, "cameras.bin"))
    write_images_bin(reconstruction.images, os.path.join(path, "images.bin"))
    write_points3D_bin(reconstruction.points3D, os.path.join(path, "points3D.bin"))
