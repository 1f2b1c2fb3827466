"""Generate tests/golden/* from the UNMODIFIED reference (oracle/_ref, built from /root/reference).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
Outputs (committed):
  kat.json            known-answer table: input recipe -> output length + SHA-256 (reference bytes)
  flat128_tile.bin    the 140-byte frame of a flat mid-grey tile at (0,0) of a larger image
  synth_40x24.jxl     a complete tiny codestream (one partial tile)
  synth_300x260.jxl   a complete 4-tile codestream with three partial tiles
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from hydrium_b200.encoder import encode_cli_loop  # noqa: E402
from hydrium_b200.synth import synth_image  # noqa: E402

sys.path.insert(0, os.path.dirname(HERE))
from util import kat_image  # noqa: E402
from oracle.pyoracle import ref_library  # noqa: E402

# name, width, height, bits, linear, shift, smooth, seed
CASES = [
    ("A_256_oneframe", 256, 256, 8, 0, -1, False, 0),
    ("B_256_tile", 256, 256, 8, 0, 0, False, 0),
    ("C_700x600", 700, 600, 8, 0, 0, False, 0),
    ("D_1920x1080", 1920, 1080, 8, 0, 0, False, 0),
    ("E_512_u16_linear", 512, 512, 16, 1, 0, False, 0),
    ("F_1024", 1024, 1024, 8, 0, 0, False, 0),
    ("S_777x333_smooth", 777, 333, 8, 0, 0, True, 3),
    ("U_515x258_u16_srgb", 515, 258, 16, 0, 0, False, 7),
    ("L_260x300_u8_linear", 260, 300, 8, 1, 0, False, 2),
    ("T_40x24", 40, 24, 8, 0, 0, False, 0),
    ("Q_300x260", 300, 260, 8, 0, 0, False, 1),
    # bits = 32: HYD_FLOAT32 samples, the 16-bit synthetic divided by 65535 in float32 (tests/util.py kat_image)
    ("P_300x260_f32_srgb", 300, 260, 32, 0, 0, False, 5),
    ("R_270x130_f32_linear", 270, 130, 32, 1, 0, False, 6),
    ("G_2304x2100_oneframe_ref_only", 2304, 2100, 8, 0, -1, False, 0),
    ("H_1024_tile512_ref_only", 1024, 1024, 8, 0, 1, False, 0),
]


def image_header_len(w, h):
    """Byte length of the image header (reference: encoder.c:164-239; SizeHeader U32 selectors
    encoder.c:98-101), without the level-10 container prefix."""
    def u32_bits(v):
        for nb in (9, 13, 18, 30):
            if v - 1 < (1 << nb):
                return nb + 2
        raise ValueError(v)
    return (33 + u32_bits(h) + u32_bits(w) + 7) // 8


def main():
    lib = ref_library("Os")
    table = {}
    for name, w, h, bits, lin, shift, smooth, seed in CASES:
        img = kat_image({"width": w, "height": h, "bits": bits, "seed": seed, "smooth": smooth})
        out = encode_cli_loop(lib, img, linear_light=lin, shift_x=shift, shift_y=shift)
        table[name] = {"width": w, "height": h, "bits": bits, "linear_light": lin, "shift": shift,
                       "smooth": smooth, "seed": seed,
                       "input_sha256": hashlib.sha256(img.tobytes()).hexdigest(),
                       "length": len(out), "sha256": hashlib.sha256(out).hexdigest()}
        if name == "T_40x24":
            open(os.path.join(HERE, "synth_40x24.jxl"), "wb").write(out)
        if name == "Q_300x260":
            open(os.path.join(HERE, "synth_300x260.jxl"), "wb").write(out)
        print(name, len(out), table[name]["sha256"][:16])
    # flat grey tile at (0,0) of a 512x512 image: header (7 bytes) + 140-byte non-last cropped frame
    img = np.full((512, 512, 3), 128, np.uint8)
    per = []
    encode_cli_loop(lib, img, tiles=[(0, 0)], is_last=0, per_tile=per)
    frame = per[0][image_header_len(512, 512):]
    open(os.path.join(HERE, "flat128_tile.bin"), "wb").write(frame)
    table["flat128_tile"] = {"length": len(frame), "sha256": hashlib.sha256(frame).hexdigest()}
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
