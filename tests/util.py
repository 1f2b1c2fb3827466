"""Helpers shared by the tests."""
import hashlib
import json
import os

import numpy as np

from hydrium_b200.synth import synth_image

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def kat_table():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


def float_image(width, height, seed=0, smooth=False):
    """HYD_FLOAT32 test image: the 16-bit synthetic scaled to [0, 1] with one float32 division."""
    return synth_image(width, height, 16, seed=seed, smooth=smooth).astype(np.float32) / np.float32(65535)


def kat_image(entry):
    if entry["bits"] == 32:
        return float_image(entry["width"], entry["height"], seed=entry["seed"], smooth=entry["smooth"])
    return synth_image(entry["width"], entry["height"], entry["bits"], seed=entry["seed"], smooth=entry["smooth"])


def sha256(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def bits_of_words(words: np.ndarray, nbits: int) -> np.ndarray:
    return np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")[:nbits]


def bits_of_bytes(b: bytes, nbits: int) -> np.ndarray:
    return np.unpackbits(np.frombuffer(b, np.uint8), bitorder="little")[:nbits]


def image_set(rng):
    """A spread of tile contents: noisy / smooth gradients, flat, saturated noise, 16-bit, edges."""
    return [
        ("synth700x600", synth_image(700, 600, 8), 0),
        ("synth512_u16_linear", synth_image(512, 512, 16), 1),
        ("flat128", np.full((256, 256, 3), 128, np.uint8), 0),
        ("black", np.zeros((264, 40, 3), np.uint8), 0),
        ("white_u16", np.full((100, 300, 3), 65535, np.uint16), 0),
        ("binary_noise", (rng.integers(0, 2, (300, 260, 3)) * 255).astype(np.uint8), 0),
        ("uniform_noise_u16", rng.integers(0, 65536, (130, 270, 3)).astype(np.uint16), 1),
        ("smooth1000x300", synth_image(1000, 300, 8, smooth=True), 0),
        ("tiny1x1", synth_image(1, 1, 8), 0),
        ("thin257x3", synth_image(257, 3, 8, seed=4), 0),
        ("float_srgb300x200", float_image(300, 200, seed=9), 0),
        ("float_linear_noise", rng.random((90, 260, 3), dtype=np.float32), 1),
        ("float_over_range", rng.random((40, 70, 3), dtype=np.float32) * np.float32(3.5), 1),
    ]
