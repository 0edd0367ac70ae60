"""CPU-side checks of the C-ABI library: it builds, loads, exports every symbol the header
declares, validates configs, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    import pcc_rl_b200
    return pcc_rl_b200._lib


def test_library_exports_every_header_symbol():
    lib = _lib()
    L = lib.load()
    hdr = open(os.path.join(ROOT, "include", "pcc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pcc_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.pcc_abi_version() == lib.PCC_ABI_VERSION == 2


def test_config_struct_layout_and_defaults():
    lib = _lib()
    L = lib.load()
    cfg = lib.PccConfig()
    L.pcc_default_config(C.byref(cfg))
    assert C.sizeof(lib.PccConfig) == 128 and C.sizeof(lib.PccConsts) == 40
    assert (cfg.history_len, cfg.n_features, list(cfg.feature_ids)[:3]) == (10, 3, [7, 10, 11])
    c = cfg.consts
    assert (c.max_rate, c.min_rate, c.delta_scale, c.reward_scale, c.max_steps, c.bytes_per_packet) == \
        (1000.0, 40.0, 0.025, 0.001, 400, 1500)
    # default ranges of the reference: bw >= 100, delay <= 0.5, queue <= 2981 packets
    assert L.pcc_ring_capacity_for(1000.0, 100.0, 0.5, 2981.0) == 65536


def test_workspace_bytes_and_validation():
    lib = _lib()
    L = lib.load()
    cfg = lib.PccConfig()
    L.pcc_default_config(C.byref(cfg))
    cfg.n_envs, cfg.ring_capacity = 4096, 65536
    sb, rb = C.c_uint64(), C.c_uint64()
    assert L.pcc_workspace_bytes(C.byref(cfg), C.byref(sb), C.byref(rb)) == 0
    assert rb.value == 4096 * 65536 * 16
    assert sb.value >= 4096 * (15 * 8 + 2 * 8 + 4 * 4 + 30 * 8)
    for field, bad in (("ring_capacity", 1000), ("history_len", 0), ("n_features", 13), ("n_envs", 0),
                       ("rng_kind", 7), ("abi_version", 99)):
        c2 = lib.PccConfig.from_buffer_copy(cfg)
        setattr(c2, field, bad)
        assert L.pcc_workspace_bytes(C.byref(c2), C.byref(sb), C.byref(rb)) == lib.PCC_EINVAL, field
        assert L.pcc_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib()
    L = lib.load()
    cfg = lib.PccConfig()
    L.pcc_default_config(C.byref(cfg))
    cfg.n_envs, cfg.ring_capacity = 4, 1024
    h = C.c_void_p()
    buf = (C.c_char * 65536)()
    addr = (C.addressof(buf) + 255) & ~255
    assert L.pcc_create(C.byref(h), C.byref(cfg), C.c_void_p(addr), C.c_void_p(addr)) == lib.PCC_ENODEV
    import pcc_rl_b200
    with pytest.raises(RuntimeError):
        pcc_rl_b200.PccBatchEnv(4)


def test_host_param_sampler_is_shard_invariant_and_in_range():
    import numpy as np
    from pcc_rl_b200 import sample_link_params
    whole = sample_link_params(5, 2, np.arange(1000), 1000)
    part = sample_link_params(5, 2, np.arange(400, 700), 1000)
    for k in whole:
        assert np.array_equal(whole[k][400:700], part[k])
    assert whole["bw"].min() >= 100 and whole["bw"].max() <= 500
    assert whole["queue"].min() >= 2 and whole["queue"].max() <= 2981
    assert (whole["start_rate"] >= 0.3 * whole["bw"]).all() and (whole["start_rate"] <= 1.5 * whole["bw"]).all()


def test_sender_obs_metadata_matches_reference_table():
    from pcc_rl_b200 import sender_obs
    assert sender_obs.feature_ids(sender_obs.DEFAULT_FEATURES) == [7, 10, 11]
    assert sender_obs.get_min_obs_vector(sender_obs.DEFAULT_FEATURES).tolist() == [-1.0, 1.0, 0.0]
    assert sender_obs.get_max_obs_vector(sender_obs.DEFAULT_FEATURES).tolist() == [10.0, 10000.0, 1000.0]
    with pytest.raises(KeyError):
        sender_obs.feature_ids("no such metric")
