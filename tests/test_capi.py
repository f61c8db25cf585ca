"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/*.h declares, refuses
to run without a device (no CPU fallback), and its host-side C matches the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import berserkerquake2_b200 as bq2
from berserkerquake2_b200 import _capi, host, synth
from scenes import bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bq2_[a-z0-9_]+)\s*\(", txt)))


@pytest.mark.parametrize("header", ["bq2_gpu.h", "bq2_synth.h"])
def test_every_declared_symbol_is_exported(header):
    names = declared(header)
    assert len(names) >= 9
    for n in names:
        assert hasattr(_capi.lib, n), f"{n} declared in include/{header} but not exported by libbq2gpu.so"
    bound = {f.__name__ for f in _capi.EXPORTS_GPU_H + _capi.EXPORTS_SYNTH_H}
    assert set(names) <= bound, set(names) - bound          # and the Python mirror binds each of them


def test_abi_version_and_record_sizes():
    assert _capi.lib.bq2_abi_version() == 2
    assert bq2.ENTITY_DTYPE.itemsize == 64 and bq2.PAIR_DTYPE.itemsize == 32
    assert bq2.WORLD_ENTITY_DTYPE.itemsize == 56 and bq2.WORLD_LIGHT_DTYPE.itemsize == 16


def test_oracle_is_not_reachable_from_the_product():
    """Nothing shipped links or imports the oracle."""
    pkg = os.path.join(ROOT, "berserkerquake2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower() or f == "_never_", os.path.join(dirpath, f)
    import subprocess
    deps = subprocess.run(["ldd", _capi.lib_path], capture_output=True, text=True).stdout
    assert "oracle" not in deps and "bq2ref" not in deps


def test_no_device_no_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(bq2.Bq2Error) as e:
        bq2.Context(0)
    assert e.value.status == -6 and "no CPU path" in str(e.value)      # BQ2_E_NODEVICE


def test_null_arguments_return_errors_not_crashes():
    L = _capi.lib
    assert L.bq2_create(0, None) == -1
    assert L.bq2_frame(None, None, None) == -1
    assert L.bq2_frame_replay(None) == -1
    assert L.bq2_download(None, 0, 0, 0, None) == -1
    assert L.bq2_upload_md2(None, None, 0, None, None, None, None, 0, None) == -1
    assert L.bq2_kernel_launches(None) == 0 and L.bq2_device(None) == -1
    L.bq2_destroy(None)


def test_host_entity_constants_match_oracle(oracle):
    """bq2_host_entity_md2 / _md3 / point_to_model are the reference's host lines (M:63526-63553,
    63825-63856, 63733-63745); the oracle restates the same lines independently."""
    blob = synth.md2(8, 5, 6, 3)
    rng = np.random.RandomState(5)
    ft = rng.randn(6, 3).astype(np.float32)
    for _ in range(40):
        f, of = int(rng.randint(6)), int(rng.randint(6))
        bl = float(np.float32(rng.rand()))
        org, oorg = (rng.randn(3) * 50).astype(np.float32), (rng.randn(3) * 50).astype(np.float32)
        ang = np.zeros(3, np.float32) if rng.rand() < 0.3 else rng.uniform(-180, 180, 3).astype(np.float32)
        e = host.entity_md2(blob, f, of, bl, org, oorg, ang)
        move, fv, bv, fl = oracle.md2_lerp_consts(blob, f, of, bl, org, oorg, ang)
        assert np.array_equal(bits(e["move"]), bits(move)) and np.array_equal(bits(e["frontv"]), bits(fv))
        assert np.array_equal(bits(e["backv"]), bits(bv)) and np.float32(e["frontlerp"]) == np.float32(fl)
        e3 = host.entity_md3(ft, f, of, bl, org, oorg, ang)
        move3, fl3 = oracle.md3_lerp_consts(ft[f], ft[of], bl, org, oorg, ang)
        assert np.array_equal(bits(e3["move"]), bits(move3)) and np.float32(e3["frontlerp"]) == np.float32(fl3)
        assert (e3["frontv"] == np.float32(fl3)).all() and (e3["backv"] == np.float32(bl)).all()
        w = (rng.randn(3) * 200).astype(np.float32)
        assert np.array_equal(bits(host.point_to_model(w, org, ang)), bits(oracle.point_to_model(w, org, ang)))


def test_casts_shadow_filter():
    assert host.casts_shadow(0)
    for f in (0x4, 0x8, 0x20, 0x80, 0x400, 0x800, 0x1000):
        assert not host.casts_shadow(f)
    assert not host.casts_shadow(0x2)                                   # RF_VIEWERMODEL, r_playershadow 0
    assert host.casts_shadow(0x2, r_playershadow=1.0, r_shadows=2.0)
    assert host.casts_shadow(0x2, r_mirror=True)


def test_shard_entities_partition():
    for n in (0, 1, 7, 8, 1024, 65536, 65537):
        for world in (1, 2, 3, 8):
            got = [host.shard_entities(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == n
            for (f0, c0), (f1, _) in zip(got, got[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in got) - min(c for _, c in got) <= 1


def test_synth_models_are_valid_md2():
    blob = synth.md2()
    h = synth.md2_header(blob)
    assert h["ident"] == 844121161 and h["version"] == 8 and h["num_xyz"] == 258 and h["num_tris"] == 512
    assert h["framesize"] == 40 + 4 * 258 and h["ofs_end"] == len(blob)
    tris = synth.md2_tris(blob)
    assert tris[:, :3].min() == 0 and tris[:, :3].max() == 257
    nb = host.md2_neighbours(blob)
    assert (nb >= 0).all()                                              # closed 2-manifold: no open edge
    for t in range(0, 512, 37):                                         # neighbour relation is symmetric
        for j in range(3):
            assert t in nb[nb[t, j]]
    assert np.array_equal(synth.md2(phase_seed=3), synth.md2(phase_seed=3))
    assert not np.array_equal(synth.md2(phase_seed=3), synth.md2(phase_seed=4))
