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
