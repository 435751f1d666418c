"""Synthetic inputs of BASELINE.json's configs: deterministic audio clips and a
background image, for benchmarks and examples (the reference's example assets are network downloads,
demo.py:29-37). The oracle keeps its own copies of the same generators; tests assert they agree."""
from __future__ import annotations

import numpy as np


def sine(seconds: float = 1.0, hz: float = 440.0, sr: int = 44100) -> np.ndarray:
    t = np.arange(int(seconds*sr))/sr
    s = np.sin(2*np.pi*hz*t).astype(np.float32)
    return np.stack([s, s])


def noise(seconds: float = 10.0, sr: int = 44100, seed: int = 0) -> np.ndarray:
    return np.random.default_rng(seed).uniform(-1, 1, (2, int(sr*seconds))).astype(np.float32)


def chirp(seconds: float = 60.0, sr: int = 44100, f0: float = 20.0, f1: float = 20000.0) -> np.ndarray:
    """L = logarithmic chirp f0→f1, R = time-reversed L, float32 (2, samples)"""
    t = np.arange(int(seconds*sr), dtype=np.float64)/sr
    beta = seconds/np.log(f1/f0)
    left = np.cos(2*np.pi*beta*f0*(np.power(f1/f0, t/seconds) - 1.0)).astype(np.float32)
    return np.stack([left, left[::-1].copy()])


def background(width: int = 1920, height: int = 1080, seed: int = 1) -> np.ndarray:
    """RGB8 image, top row first: smooth colour field + grain"""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:height, 0:width].astype(np.float64)
    x /= width; y /= height
    img = np.stack([0.5 + 0.5*np.sin(6.0*x + 2.0*np.cos(5.0*y)),
                    0.5 + 0.5*np.sin(4.0*y + 3.0*x*x + 1.0),
                    0.5 + 0.5*np.cos(9.0*(x - 0.5)*(y - 0.5) + 2.0)], -1)
    img += rng.uniform(-0.08, 0.08, img.shape)
    return (np.clip(img, 0, 1)*255).round().astype(np.uint8)


def video_frames(width: int = 320, height: int = 180, frames: int = 30, seed: int = 3) -> np.ndarray:
    """Deterministic rgb24 test clip (frames, height, width, 3), top row first: the background drifting sideways with a
    frame counter stripe, so every frame and both orientations are distinguishable"""
    base = background(width, height, seed)
    out = np.empty((frames, height, width, 3), np.uint8)
    for k in range(frames):
        out[k] = np.roll(base, 7*k, axis=1)
        out[k, :max(2, height//16), :(k + 1)*width//max(frames, 1)] = (255, 255 - 8*k % 256, 8*k % 256)
    return out


def write_y4m(path, rgb: np.ndarray, fps: int = 30, colorspace: str = "420jpeg") -> None:
    """rgb24 frames (frames, H, W, 3, top row first) → a YUV4MPEG2 file: BT.601 limited range, chroma = box mean of the
    samples it covers (even W and H for 420 / 422). The inverse of what sfb_video_frame undoes, up to the chroma loss."""
    rgb = np.asarray(rgb, np.float32)
    r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    y = 16.0 + 0.256788*r + 0.504129*g + 0.097906*b
    u = 128.0 - 0.148223*r - 0.290993*g + 0.439216*b
    v = 128.0 + 0.439216*r - 0.367788*g - 0.071427*b
    if colorspace.startswith("420"):
        u = u.reshape(u.shape[0], u.shape[1]//2, 2, u.shape[2]//2, 2).mean(axis=(2, 4))
        v = v.reshape(v.shape[0], v.shape[1]//2, 2, v.shape[2]//2, 2).mean(axis=(2, 4))
    elif colorspace.startswith("422"):
        u = u.reshape(u.shape[0], u.shape[1], u.shape[2]//2, 2).mean(axis=3)
        v = v.reshape(v.shape[0], v.shape[1], v.shape[2]//2, 2).mean(axis=3)
    planes = [np.clip(np.rint(p), 0, 255).astype(np.uint8) for p in (y, u, v)]
    with open(path, "wb") as f:
        f.write(f"YUV4MPEG2 W{rgb.shape[2]} H{rgb.shape[1]} F{fps}:1 Ip A1:1 C{colorspace}\n".encode())
        for k in range(rgb.shape[0]):
            f.write(b"FRAME\n")
            for p in planes:
                f.write(p[k].tobytes())
