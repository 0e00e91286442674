"""CPU-side checks of the boundary: the shared library loads and exports every symbol the header declares."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from dibs_b200 import _native
    if not os.path.exists(_native.LIB_PATH):
        from dibs_b200.build import build
        build()
    return _native


def test_header_symbols_exported_and_bound():
    nat = _lib()
    hdr = open(os.path.join(ROOT, "include", "dibs_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:const\s+char\*|int64_t|int)\s+(dibs_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 20
    handle = ctypes.CDLL(nat.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/dibs_b200.h but not exported"
    assert declared == set(nat.PROTOTYPES), declared ^ set(nat.PROTOTYPES)
    assert nat.lib().dibs_abi_version() == 1


def test_config_struct_layout_matches_header():
    nat = _lib()
    hdr = open(os.path.join(ROOT, "include", "dibs_b200.h")).read()
    body = hdr[hdr.index("typedef struct dibs_config {"):hdr.index("} dibs_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in re.findall(r"(?:int32_t|float)\s+([^;]+);", body):
        fields += [f.strip() for f in decl.split(",")]
    assert fields == [f[0] for f in nat.DibsConfig._fields_]
    assert ctypes.sizeof(nat.DibsConfig) == 4 * len(fields)


def test_host_prng_split_matches_oracle():
    from dibs_b200.inference.dibs import split, PRNGKey
    from oracle import threefry as tf
    for seed in (0, 123, 987654321987):
        for num in (2, 5, 257):
            assert (split(PRNGKey(seed), num) == tf.split(tf.prng_key(seed), num)).all()
            assert (split(PRNGKey(seed), num, True) == tf.split(tf.prng_key(seed), num, partitionable=True)).all()


def test_invalid_configs_map_to_reference_exceptions():
    nat = _lib()
    c = nat.DibsConfig()
    out = ctypes.c_void_p()
    assert nat.lib().dibs_plan_create(ctypes.byref(c), ctypes.byref(out)) == -1     # ValueError
    with pytest.raises(ValueError):
        nat.check(-1)
    with pytest.raises(NotImplementedError):
        nat.check(-2)


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (no CPU fallback)."""
    pkg = os.path.join(ROOT, "dibs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
