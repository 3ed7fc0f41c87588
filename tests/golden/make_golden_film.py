#!/usr/bin/env python
"""Generate tests/golden/film_export.npz from the UNMODIFIED reference (oracle/_ref): Film::to_byte_array
(camera.cc:27-48) and the file stbi_write_hdr writes for Film::to_float_array (main.cc:125-126,
stb_image_write.h:601-740) for the film of tests.common.export_test_film, plus a narrow film (nx < 8: the
writer's flat, un-encoded scanlines).  Square films only: the reference's Film indexes y*ny + x (camera.cc:12-15).  Run in the build container:  python tests/golden/make_golden_film.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.bindings import Ref  # noqa: E402
from tests.common import export_test_film  # noqa: E402

ref = Ref()
wide = export_test_film(64, 64, 3)
big = export_test_film(160, 160, 4)   # rows longer than the 127 / 128 byte pieces of the RLE
narrow = export_test_film(7, 7, 5)
np.savez_compressed(os.path.join(HERE, "film_export.npz"),
                    wide=wide, wide_rgb8=ref.film_to_bytes(wide), wide_hdr=np.frombuffer(ref.write_hdr(wide), np.uint8),
                    big=big, big_rgb8=ref.film_to_bytes(big), big_hdr=np.frombuffer(ref.write_hdr(big), np.uint8),
                    narrow=narrow, narrow_rgb8=ref.film_to_bytes(narrow),
                    narrow_hdr=np.frombuffer(ref.write_hdr(narrow), np.uint8))
print("wrote film_export.npz")
