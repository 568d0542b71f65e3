"""Shared fixtures.  Tests marked `gpu` call the CUDA path through the C ABI (libb200rt.so) and
check it against the oracle; everything else runs on CPU (oracle vs golden vectors, host logic,
C-ABI export table, gloo world_size-2 logic)."""
import gzip
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

SMALL_SCENES = ["rtow_final", "rtow_lights", "quads", "cornell_empty", "cornell", "xmas", "pathological"]
RENDER_SCENES = ["rtow_final", "rtow_lights", "quads", "cornell_empty", "cornell", "xmas"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with -m gpu on the GPU box.  If they get collected on a machine
    # without a device (plain `pytest tests/`), skip them instead of erroring -- but a MISSING
    # LIBRARY is never skipped: importing cpp_raytracer_b200.capi.lib() fails loudly.
    from cpp_raytracer_b200 import capi
    if capi.device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    from cpp_raytracer_b200 import scene_io

    class G:
        root = GOLDEN

        @staticmethod
        def scene(name):
            return scene_io.load_scene(os.path.join(GOLDEN, f"{name}.scene.gz"))

        @staticmethod
        def rays(name):
            return scene_io.load_rays(os.path.join(GOLDEN, f"{name}.rays.gz"))

        @staticmethod
        def hits(name, brute=False):
            return scene_io.load_hits(os.path.join(GOLDEN, f"{name}.hits_brute.gz" if brute else f"{name}.hits.gz"))

        @staticmethod
        def ref_image(name, tag):
            with gzip.open(os.path.join(GOLDEN, f"{name}.{tag}.npy.gz"), "rb") as f:
                return np.load(f)

        @staticmethod
        def summary():
            with open(os.path.join(GOLDEN, "golden_summary.json")) as f:
                return json.load(f)

        @staticmethod
        def kat():
            with open(os.path.join(GOLDEN, "kat.json")) as f:
                return json.load(f)

    return G


@pytest.fixture(scope="session")
def ref_bridge():
    """Path of the compiled reference (oracle/_ref/ref_bridge), or None if it was not built."""
    p = os.path.join(ROOT, "oracle", "_ref", "ref_bridge")
    return p if os.path.exists(p) else None
