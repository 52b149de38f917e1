"""Host-side mirror of the reference camera (Python plumbing for bench.py / tests; the C++ mirror is host/VoxelCamera.h).

Same names and argument meaning as Runtimes/Instance/VoxelCamera.{h,cpp}:
  FVoxelCamera.InitializeVoxelCamera / GetProjectionMatrix / GetViewMatrix / GetForwardVector / GetCameraUniform /
  UpdateCamera, and FVoxelMathHelper::ConvertToChunkLocation (Runtimes/Helper/VoxelMathHelper.h:17-22).
glm 0.9.9.8 perspectiveRH_ZO / lookAtRH and the Cookbook CameraPositioner_FirstPerson view matrix are restated from their
published definitions in fp32.  The kernels consume the resulting explicit matrices (FGPUUniformCamera, 160 B).
"""
import math

import numpy as np

from .capi import Camera

f32 = np.float32

try:  # glm's tan(float) is libm tanf; use the same routine so the matrices match a C++ host bit for bit
    import ctypes as _C
    import ctypes.util as _CU
    _libm = _C.CDLL(_CU.find_library("m") or "libm.so.6")
    _libm.tanf.restype = _C.c_float
    _libm.tanf.argtypes = [_C.c_float]

    def _tanf(x):
        return f32(_libm.tanf(float(x)))
except Exception:  # pragma: no cover
    def _tanf(x):
        return f32(np.tan(f32(x)))


def perspective_rh_zo(fovy, aspect, z_near, z_far):
    """glm::perspective under GLM_FORCE_DEPTH_ZERO_TO_ONE, right-handed (reference CMakeLists.txt:5)."""
    m = np.zeros((4, 4), dtype=f32)  # m[col][row]
    t = _tanf(f32(fovy) / f32(2.0))
    m[0][0] = f32(1.0) / (f32(aspect) * t)
    m[1][1] = f32(1.0) / t
    m[2][2] = f32(z_far) / (f32(z_near) - f32(z_far))
    m[2][3] = f32(-1.0)
    m[3][2] = -(f32(z_far) * f32(z_near)) / (f32(z_far) - f32(z_near))
    return m


def _normalize(v):
    inv = f32(1.0) / np.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2], dtype=f32)
    return np.array([v[0] * inv, v[1] * inv, v[2] * inv], dtype=f32)


def _cross(x, y):
    return np.array([x[1] * y[2] - y[1] * x[2], x[2] * y[0] - y[2] * x[0], x[0] * y[1] - y[0] * x[1]], dtype=f32)


def look_at_rotation(eye, center, up):
    eye, center, up = (np.asarray(a, dtype=f32) for a in (eye, center, up))
    f = _normalize(center - eye)
    s = _normalize(_cross(f, up))
    u = _cross(s, f)
    m = np.zeros((4, 4), dtype=f32)
    for c in range(3):
        m[c][0] = s[c]
        m[c][1] = u[c]
        m[c][2] = -f[c]
    m[3][3] = f32(1.0)
    return m


def convert_to_chunk_location(position, chunk_size):
    """FVoxelMathHelper::ConvertToChunkLocation -> (fracted position, chunk offset)."""
    p = np.asarray(position, dtype=f32)
    c = np.floor(p / f32(chunk_size)).astype(f32)
    return (p - c * f32(chunk_size)).astype(f32), c.astype(np.int32)


class FVoxelCamera:
    def __init__(self, position=(5.0, 2.0, 2.0), target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0)):  # VoxelCamera.h:23
        self.Position = np.asarray(position, dtype=f32)
        self.Orientation = look_at_rotation(position, target, up)
        self.Fov = f32(45.0 * (math.pi / 180.0))
        self.Near = f32(0.1)
        self.Far = f32(1000.0)
        self.bReverseZ = True
        self.CameraChunkLocation = np.zeros(3, dtype=np.int32)
        self.CameraForward = np.zeros(3, dtype=f32)
        self.CameraChunkUpdateCallback = None
        self.CameraUpdateCallback = None

    def InitializeVoxelCamera(self, FovAngle=45.0, Near=0.5, Far=1000.0, bReverseZ=True):
        self.Fov = f32(float(FovAngle) * (math.pi / float(f32(180.0))))
        self.bReverseZ = bool(bReverseZ)
        self.Near = f32(Near)
        self.Far = f32(Far)

    def GetProjectionMatrix(self, ViewWidth=1920.0, ViewHeight=1280.0):
        aspect = f32(ViewWidth) / f32(ViewHeight)
        if self.bReverseZ:
            return perspective_rh_zo(self.Fov, aspect, self.Far, self.Near)
        return perspective_rh_zo(self.Fov, aspect, self.Near, self.Far)

    def GetViewMatrix(self):
        m = self.Orientation.copy()
        t = -self.Position
        for r in range(3):
            m[3][r] = ((m[0][r] * t[0] + m[1][r] * t[1]) + m[2][r] * t[2]) + f32(0.0)
        return m

    def GetForwardVector(self):
        v = self.GetViewMatrix()
        return -np.array([v[0][2], v[1][2], v[2][2]], dtype=f32)

    def UpdateCamera(self, chunk_size=16.0):
        fr, off = convert_to_chunk_location(self.Position, chunk_size)
        if np.any(off != 0):
            if self.CameraChunkUpdateCallback:
                self.CameraChunkUpdateCallback()
            self.CameraChunkLocation = self.CameraChunkLocation + off
        fwd = self.GetForwardVector()
        if np.any(fwd != self.CameraForward) or np.any(off != 0):
            self.CameraForward = fwd
            if self.CameraUpdateCallback:
                self.CameraUpdateCallback()
        self.Position = fr

    def GetCameraUniform(self, ViewWidth=1920.0, ViewHeight=1280.0):
        cam = np.zeros(1, dtype=Camera)
        cam["Projection"][0] = self.GetProjectionMatrix(ViewWidth, ViewHeight).reshape(16)
        cam["View"][0] = self.GetViewMatrix().reshape(16)
        cam["CameraChunkLocation"][0][:3] = self.CameraChunkLocation
        cam["SubCameraLocation"][0][:3] = np.trunc(self.Position).astype(f32)  # ivec4(getPosition(), 0)  VoxelCamera.cpp:38
        return cam


def camera_uniform(eye, center, width, height, up=(0.0, 0.0, 1.0), fov_deg=60.0, z_near=0.1, z_far=1000.0, chunk_size=16.0):
    """The FGPUUniformCamera the frame loop uploads for a camera placed at `eye` looking at `center`
    (defaults: VoxelWindowsInstance.h:22-24)."""
    c = FVoxelCamera(eye, center, up)
    c.InitializeVoxelCamera(fov_deg, z_near, z_far, True)
    c.UpdateCamera(chunk_size)
    return c.GetCameraUniform(float(width), float(height))
