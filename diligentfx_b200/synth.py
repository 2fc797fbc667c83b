"""Seeded synthetic G-buffer generator (SURVEY.md §8d "Synthetic G-buffer").

Produces, for a sequence of frames, the inputs the reference chain consumes
(Hydrogent/src/Tasks/HnPostProcessTask.cpp:776-832): depth, normal, colour, material, motion vectors, previous depth
and the curr/prev `CameraAttribs`. The scene is an analytic ray-cast (ground plane + K seeded spheres + sky), so any
resolution can be generated and depth / normals / motion are mutually consistent with the cameras:

* camera: LH perspective, fovY 60 deg, near 0.1, far 100 (D3D convention, row-vector matrices), translated a little
  every frame; the projection is jittered exactly like the reference does it
  (TemporalAntiAliasing::GetJitteredProjMatrix, TemporalAntiAliasing.hpp:138-156; Hydrogent/src/HnCamera.cpp:124-139;
  camera block filled like Hydrogent/src/Tasks/HnBeginFrameTask.cpp:490-515: mProj/mViewProj are the jittered ones).
* motion = (NDC_curr - jitter_curr) - (NDC_prev - jitter_prev)  (Shaders/Common/public/ShaderUtilities.fxh:88-91).
* depth 1.0 = background (sky); colour is HDR with log-uniform highlights so the Bloom threshold and the TAA
  clipping are exercised; material.x = roughness from {0, 0.05, 0.15, 0.3, 0.6} so about half the pixels pass
  the SSR roughness threshold.

Everything is numpy on the host (it feeds both the CUDA path and the CPU oracle with identical bits).
"""
from __future__ import annotations

import ctypes
import math
import os
from dataclasses import dataclass, field

import numpy as np

from .capi import CameraAttribs

NEAR, FAR, FOVY = 0.1, 100.0, math.radians(60.0)
PITCH = math.radians(18.0)


def halton(base: int, index: int) -> float:
    """TemporalAntiAliasing.cpp:43-54 (host-side float32 arithmetic)."""
    result = np.float32(0.0)
    f = np.float32(1.0)
    while index > 0:
        f = np.float32(f / np.float32(base))
        result = np.float32(result + f * np.float32(index % base))
        index = int(math.floor(np.float32(index) / np.float32(base)))
    return float(result)


def taa_jitter(frame: int, width: int, height: int) -> tuple[float, float]:
    """TemporalAntiAliasing::GetJitterOffset, TemporalAntiAliasing.cpp:63-78."""
    jx = np.float32(np.float32(halton(2, (frame % 16) + 1)) - np.float32(0.5)) / np.float32(np.float32(0.5) * np.float32(width))
    jy = np.float32(np.float32(halton(3, (frame % 16) + 1)) - np.float32(0.5)) / np.float32(np.float32(0.5) * np.float32(height))
    return float(jx), float(jy)


def _proj(width: int, height: int, jitter: tuple[float, float]) -> np.ndarray:
    aspect = width / height
    t = math.tan(FOVY / 2)
    m = np.zeros((4, 4), np.float64)
    m[0, 0] = 1.0 / (aspect * t)
    m[1, 1] = 1.0 / t
    m[2, 2] = FAR / (FAR - NEAR)
    m[2, 3] = 1.0
    m[3, 2] = -NEAR * FAR / (FAR - NEAR)
    # GetJitteredProjMatrix: perspective -> m20 += jx, m21 += jy
    m[2, 0] += jitter[0]
    m[2, 1] += jitter[1]
    return m


def _view(pos: np.ndarray, yaw: float) -> tuple[np.ndarray, np.ndarray]:
    """Returns (view, world) row-vector matrices for a camera at `pos` rotated by `yaw` about +y."""
    c, s = math.cos(yaw), math.sin(yaw)
    cp, sp = math.cos(PITCH), math.sin(PITCH)
    yaw_m = np.array([[c, 0, -s, 0], [0, 1, 0, 0], [s, 0, c, 0], [0, 0, 0, 1]], np.float64)
    pitch_m = np.array([[1, 0, 0, 0], [0, cp, sp, 0], [0, -sp, cp, 0], [0, 0, 0, 1]], np.float64)  # look down by PITCH
    rot = pitch_m @ yaw_m                                                                        # camera->world rotation rows
    world = rot.copy()
    world[3, :3] = pos
    view = np.linalg.inv(world)
    return view, world


@dataclass
class CameraState:
    view: np.ndarray
    world: np.ndarray
    proj: np.ndarray
    jitter: tuple[float, float]
    pos: np.ndarray
    attribs: CameraAttribs = field(default=None)


def make_camera(frame: int, width: int, height: int, use_jitter: bool = True) -> CameraState:
    pos = np.array([0.004 * frame, 0.15 + 0.001 * frame, -0.002 * frame], np.float64)
    yaw = 0.0004 * frame
    jitter = taa_jitter(frame, width, height) if use_jitter else (0.0, 0.0)
    view, world = _view(pos, yaw)
    proj = _proj(width, height, jitter)
    vp = view @ proj
    cam = CameraAttribs()
    cam.f4Position[:] = [pos[0], pos[1], pos[2], 1.0]
    cam.f4ViewportSize[:] = [float(width), float(height), float(np.float32(1.0) / np.float32(width)), float(np.float32(1.0) / np.float32(height))]
    # SetClipPlanes(0.1, 100) — BasicStructures.fxh:134-147
    cam.fNearPlaneZ, cam.fFarPlaneZ, cam.fNearPlaneDepth, cam.fFarPlaneDepth = NEAR, FAR, 0.0, 1.0
    cam.fSceneNearZ, cam.fSceneFarZ, cam.fSceneNearDepth, cam.fSceneFarDepth = NEAR, FAR, 0.0, 1.0
    cam.fHandness = 1.0 if np.linalg.det(view) > 0 else -1.0
    cam.uiFrameIndex = frame
    cam.fFocusDistance, cam.fFStop, cam.fFocalLength, cam.fSensorWidth = 10.0, 5.6, 50.0, 36.0
    cam.fSensorHeight, cam.fExposure = 24.0, 0.0
    cam.f2Jitter[:] = [jitter[0], jitter[1]]

    def put(dst, m):
        m32 = np.asarray(m, np.float32)
        for r in range(4):
            for c in range(4):
                dst.m[r][c] = float(m32[r, c])

    put(cam.mView, view), put(cam.mProj, proj), put(cam.mViewProj, vp)
    put(cam.mViewInv, world), put(cam.mProjInv, np.linalg.inv(proj)), put(cam.mViewProjInv, np.linalg.inv(vp))
    return CameraState(view, world, proj, jitter, pos, cam)


@dataclass
class Scene:
    centers: np.ndarray    # (K,3)
    radii: np.ndarray      # (K,)
    albedo: np.ndarray     # (K+1,3)   last = ground
    rough: np.ndarray      # (K+1,)
    emissive: np.ndarray   # (K+1,)


def make_scene(seed: int = 7, k: int = 24) -> Scene:
    rng = np.random.default_rng(seed)
    z = rng.uniform(3.0, 28.0, k)
    x = rng.uniform(-0.55, 0.55, k) * z * 1.6
    r = rng.uniform(0.35, 1.4, k)
    y = -1.0 + r * rng.uniform(0.6, 1.3, k)
    centers = np.stack([x, y, z], 1)
    albedo = rng.uniform(0.15, 0.95, (k + 1, 3))
    albedo[-1] = [0.55, 0.55, 0.6]
    rough_set = np.array([0.0, 0.05, 0.15, 0.3, 0.6])
    rough = rough_set[rng.integers(0, 5, k + 1)]
    rough[-1] = 0.12
    emissive = np.where(rng.uniform(size=k + 1) < 0.2, np.exp2(rng.uniform(2.0, 6.0, k + 1)), 0.0)
    emissive[-1] = 0.0
    return Scene(centers, r, albedo, rough, emissive)


def _raycast(scene: Scene, cam: CameraState, width: int, height: int, rows: slice):
    """Ray-cast rows `rows` of the frame. Returns dict of per-pixel arrays (float64)."""
    ys = np.arange(rows.start, rows.stop, dtype=np.float64)
    xs = np.arange(width, dtype=np.float64)
    ndc_x = (2.0 * (xs + 0.5) / width - 1.0)[None, :]
    ndc_y = (1.0 - 2.0 * (ys + 0.5) / height)[:, None]
    p = cam.proj
    # clip = v * P (row vector): ndc.x = (x*m00 + z*m20)/z  ->  x/z = (ndc.x - m20)/m00
    dx = np.broadcast_to((ndc_x - p[2, 0]) / p[0, 0], (len(ys), width))
    dy = np.broadcast_to((ndc_y - p[2, 1]) / p[1, 1], (len(ys), width))
    dv = np.stack([dx, dy, np.ones_like(dx)], -1)           # view-space direction with dz == 1  => t == view z
    rot = cam.world[:3, :3]
    dw = dv @ rot                                            # world direction (row-vector)
    o = cam.pos

    t_best = np.full(dx.shape, np.inf)
    obj = np.full(dx.shape, -1, np.int32)
    # ground plane y = -1
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = (-1.0 - o[1]) / dw[..., 1]
    hit = (tg > NEAR) & (tg < FAR * 0.6) & np.isfinite(tg)
    t_best = np.where(hit, tg, t_best)
    obj = np.where(hit, len(scene.radii), obj)
    a = np.sum(dw * dw, -1)
    for k in range(len(scene.radii)):
        oc = o - scene.centers[k]
        b = 2.0 * (dw @ oc)
        c = float(oc @ oc) - scene.radii[k] ** 2
        disc = b * b - 4 * a * c
        ok = disc > 0
        sq = np.sqrt(np.where(ok, disc, 0.0))
        t0 = (-b - sq) / (2 * a)
        hit = ok & (t0 > NEAR) & (t0 < t_best)
        t_best = np.where(hit, t0, t_best)
        obj = np.where(hit, k, obj)

    is_bg = obj < 0
    t = np.where(is_bg, FAR, t_best)
    pw = o + dw * t[..., None]
    normal = np.zeros_like(pw)
    ground = obj == len(scene.radii)
    normal[ground] = [0.0, 1.0, 0.0]
    sph = (~is_bg) & (~ground)
    idx = np.clip(obj, 0, len(scene.radii) - 1)
    nrm = (pw - scene.centers[idx]) / scene.radii[idx][..., None]
    normal = np.where(sph[..., None], nrm, normal)
    normal[is_bg] = -dw[is_bg] / np.linalg.norm(dw[is_bg], axis=-1, keepdims=True) if is_bg.any() else normal[is_bg]
    return dict(t=t, pw=pw, normal=normal, obj=obj, is_bg=is_bg, ndc_x=np.broadcast_to(ndc_x, dx.shape), ndc_y=np.broadcast_to(ndc_y, dx.shape))


def _hash01(a: np.ndarray, b: np.ndarray, seed: int) -> np.ndarray:
    h = (a.astype(np.uint64) * np.uint64(73856093)) ^ (b.astype(np.uint64) * np.uint64(19349663)) ^ np.uint64(seed * 83492791)
    h = (h ^ (h >> np.uint64(13))) * np.uint64(1274126177)
    h = h ^ (h >> np.uint64(16))
    return (h & np.uint64(0xFFFFFF)).astype(np.float64) / float(0x1000000)


def generate_frame(scene: Scene, frame: int, width: int, height: int, prev_depth: np.ndarray | None = None,
                   use_jitter: bool = True, chunk_rows: int = 256) -> dict:
    """Returns dict(depth (H,W) f32, normal (H,W,4), color (H,W,4), material (H,W,4), motion (H,W,2), prev_depth (H,W),
    curr_camera, prev_camera (CameraAttribs))."""
    return generate_rows(scene, frame, width, height, 0, height, prev_depth, use_jitter, chunk_rows)


def _map_chunks(fn, starts) -> None:
    starts = list(starts)
    workers = min(len(starts), max(1, min(16, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))))
    if workers <= 1:
        for g0 in starts:
            fn(g0)
        return
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=workers) as pool:
        list(pool.map(fn, starts))


def generate_rows(scene: Scene, frame: int, width: int, height: int, row0: int, row1: int, prev_depth: np.ndarray | None = None,
                  use_jitter: bool = True, chunk_rows: int = 256) -> dict:
    """Rows [row0, row1) of frame `frame` of the full width x height image (arrays have row1 - row0 rows). When `prev_depth`
    is None the previous frame's depth of the same rows is ray-cast too (frame 0: a copy of the current depth)."""
    cam = make_camera(frame, width, height, use_jitter)
    prev = make_camera(max(frame - 1, 0), width, height, use_jitter)
    nrows = row1 - row0
    depth = np.empty((nrows, width), np.float32)
    normal = np.zeros((nrows, width, 4), np.float32)
    color = np.zeros((nrows, width, 4), np.float32)
    material = np.zeros((nrows, width, 4), np.float32)
    motion = np.empty((nrows, width, 2), np.float32)
    light = np.array([0.4, 0.8, -0.45])
    light = light / np.linalg.norm(light)
    nobj = len(scene.radii)
    pvp = prev.view @ prev.proj

    def chunk(g0: int) -> None:
        grows = slice(g0, min(g0 + chunk_rows, row1))
        rows = slice(g0 - row0, grows.stop - row0)
        rc = _raycast(scene, cam, width, height, grows)
        z = rc["t"]
        p = cam.proj
        d = (z * p[2, 2] + p[3, 2]) / z
        d = np.where(rc["is_bg"], 1.0, d)
        depth[rows] = d.astype(np.float32)
        normal[rows, :, :3] = rc["normal"].astype(np.float32)
        # motion: curr NDC of the pixel centre is (ndc_x, ndc_y) by construction
        pw1 = np.concatenate([rc["pw"], np.ones_like(z)[..., None]], -1)
        clip_prev = pw1 @ pvp
        ndc_prev = clip_prev[..., :2] / clip_prev[..., 3:4]
        mx = (rc["ndc_x"] - cam.jitter[0]) - (ndc_prev[..., 0] - prev.jitter[0])
        my = (rc["ndc_y"] - cam.jitter[1]) - (ndc_prev[..., 1] - prev.jitter[1])
        motion[rows, :, 0] = mx.astype(np.float32)
        motion[rows, :, 1] = my.astype(np.float32)
        # shading
        obj = np.where(rc["is_bg"], nobj, rc["obj"])
        oid = np.clip(obj, 0, nobj)
        alb = scene.albedo[oid]
        ndl = np.clip(rc["normal"] @ light, 0.0, 1.0)
        # checker on the ground, stripes on spheres: gives TAA / bloom some texture
        gx = np.floor(rc["pw"][..., 0] * 1.5).astype(np.int64)
        gz = np.floor(rc["pw"][..., 2] * 1.5).astype(np.int64)
        checker = np.where(((gx + gz) & 1) == 0, 1.0, 0.55)
        tex = np.where(obj == nobj, checker, 0.8 + 0.2 * np.sin(rc["pw"][..., 1] * 9.0))
        col = alb * (0.08 + 1.6 * ndl[..., None]) * tex[..., None]
        col = col + alb * scene.emissive[oid][..., None]
        # sparse log-uniform highlights (5 % of pixels in [4, 64]) keyed on world cell so they are temporally stable
        hx = np.floor(rc["pw"][..., 0] * 40.0).astype(np.int64)
        hz = np.floor((rc["pw"][..., 2] + rc["pw"][..., 1]) * 40.0).astype(np.int64)
        h = _hash01(hx, hz, 11)
        spark = np.where(h < 0.05, np.exp2(2.0 + 4.0 * _hash01(hx, hz, 29)), 0.0)
        col = col + spark[..., None] * alb
        sky = np.stack([0.35 + 0.3 * rc["ndc_y"], 0.5 + 0.3 * rc["ndc_y"], 0.9 + 0.0 * rc["ndc_y"]], -1) * 1.2
        col = np.where(rc["is_bg"][..., None], sky, col)
        color[rows, :, :3] = np.maximum(col, 0.0).astype(np.float32)
        color[rows, :, 3] = 1.0
        material[rows, :, 0] = np.where(rc["is_bg"], 1.0, scene.rough[oid]).astype(np.float32)
        material[rows, :, 1] = np.where(obj == nobj, 0.0, 0.5).astype(np.float32)

    # Chunks are independent (every value is a per-pixel function) and numpy releases the GIL inside its kernels, so they are
    # ray-cast on a thread pool: same chunk size, same arithmetic, same bits as the sequential loop — only the wall time of
    # a 4K frame changes (bench.py generates its inputs at start-up, also under torchrun where OMP_NUM_THREADS is 1).
    _map_chunks(chunk, range(row0, row1, chunk_rows))
    if prev_depth is None:
        if frame == 0:
            prev_depth = depth.copy()
        else:
            prev_depth = np.empty_like(depth)

            def prev_chunk(g0: int) -> None:
                grows = slice(g0, min(g0 + chunk_rows, row1))
                rc = _raycast(scene, prev, width, height, grows)
                z = rc["t"]
                prev_depth[g0 - row0:grows.stop - row0] = np.where(rc["is_bg"], 1.0, (z * prev.proj[2, 2] + prev.proj[3, 2]) / z).astype(np.float32)

            _map_chunks(prev_chunk, range(row0, row1, chunk_rows))
    return dict(depth=depth, normal=normal, color=color, material=material, motion=motion, prev_depth=prev_depth,
                curr_camera=cam.attribs, prev_camera=prev.attribs, frame=frame)


def generate_sequence(width: int, height: int, frames: int, seed: int = 7, k: int = 24, use_jitter: bool = True, first_frame: int = 0) -> list[dict]:
    scene = make_scene(seed, k)
    out = []
    prev_depth = None
    if first_frame > 0:
        prev_depth = generate_frame(scene, first_frame - 1, width, height, None, use_jitter)["depth"]
    for f in range(first_frame, first_frame + frames):
        fr = generate_frame(scene, f, width, height, prev_depth, use_jitter)
        prev_depth = fr["depth"]
        out.append(fr)
    return out


def noise_frame(width: int, height: int, seed: int = 3) -> dict:
    """Pure-noise variant (worst case for caches / branch coherence): white-noise depth, random normals."""
    rng = np.random.default_rng(seed)
    cam = make_camera(0, width, height, False)
    depth = rng.uniform(0.9, 0.9999, (height, width)).astype(np.float32)
    depth[rng.uniform(size=depth.shape) < 0.1] = 1.0
    n = rng.normal(size=(height, width, 3))
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    normal = np.zeros((height, width, 4), np.float32)
    normal[..., :3] = n
    color = np.zeros((height, width, 4), np.float32)
    color[..., :3] = np.exp2(rng.uniform(-8, 6, (height, width, 3)))
    color[..., 3] = 1
    material = np.zeros((height, width, 4), np.float32)
    material[..., 0] = rng.choice([0.0, 0.05, 0.15, 0.3, 0.6], (height, width))
    motion = (rng.uniform(-1, 1, (height, width, 2)) * 4.0 / np.array([width, height])).astype(np.float32)
    return dict(depth=depth, normal=normal, color=color, material=material, motion=motion, prev_depth=depth.copy(),
                curr_camera=cam.attribs, prev_camera=cam.attribs, frame=0)


def _reversed_camera(a: CameraAttribs) -> CameraAttribs:
    """The same camera with a reversed-depth projection: depth' = 1 - depth, i.e. (row-vector convention, clip.z = z*m22 + m32,
    clip.w = z) m22' = 1 - m22, m32' = -m32; the derived matrices are recomputed in float64 from the stored float32 ones."""
    def get(m):
        return np.array([[m.m[r][c] for c in range(4)] for r in range(4)], np.float64)

    def put(dst, m):
        m32 = np.asarray(m, np.float32)
        for r in range(4):
            for c in range(4):
                dst.m[r][c] = float(m32[r, c])

    out = CameraAttribs.from_buffer_copy(bytes(a))
    view, proj = get(a.mView), get(a.mProj)
    proj[2, 2], proj[3, 2] = 1.0 - proj[2, 2], -proj[3, 2]
    vp = view @ proj
    put(out.mProj, proj), put(out.mViewProj, vp)
    put(out.mProjInv, np.linalg.inv(proj)), put(out.mViewProjInv, np.linalg.inv(vp))
    out.fNearPlaneDepth, out.fFarPlaneDepth, out.fSceneNearDepth, out.fSceneFarDepth = 1.0, 0.0, 1.0, 0.0
    return out


def reverse_depth_frame(frame: dict) -> dict:
    """The same frame as a renderer with a reversed depth buffer would produce it (PostFXContext::FEATURE_FLAG_REVERSED_DEPTH):
    depth' = 1 - depth (exact in fp32 for depth >= 0.5, which is where this scene lives), background = 0, cameras with the
    reversed projection. Everything else is unchanged."""
    out = dict(frame)
    for k in ("depth", "prev_depth"):
        out[k] = (np.float32(1.0) - np.asarray(frame[k], np.float32)).astype(np.float32)
    out["curr_camera"], out["prev_camera"] = _reversed_camera(frame["curr_camera"]), _reversed_camera(frame["prev_camera"])
    return out
