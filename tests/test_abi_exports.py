"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/yavo_b200.h declares;
without a GPU the product fails loudly instead of falling back."""
import ctypes as C
import os
import re

import pytest


def declared_symbols():
    from ya_vo_b200 import capi
    txt = open(capi.HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(yavo_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_header(cuda_lib):
    names = declared_symbols()
    assert set(names) == set(cuda_lib.EXPORTS)
    L = C.CDLL(cuda_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n


def test_header_cites_reference_interfaces(cuda_lib):
    txt = open(cuda_lib.HEADER).read()
    for cite in ("src/FastDetector.cc:277-369", "src/BriefDescriptor.cc:86-124", "src/BriefDescriptor.cc:163-183",
                 "src/Image.cc:8-17", "src/BriefDescriptor.cc:213-231"):
        assert cite in txt


def test_library_is_sm100a_and_blackwell_native(cuda_lib):
    """The shipped library holds sm_100a code only, and its SASS shows what the design claims: tcgen05 tensor-core
    instructions with TMEM loads / stores (the matcher), tensor copies (TMA: the detect kernel's tile, the BRIEF kernel's
    patches) and bulk copies."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", cuda_lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out
    sass = subprocess.run(["cuobjdump", "-sass", cuda_lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCQMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP"):
        assert mnemonic in sass, mnemonic


def test_no_cpu_fallback_without_device(cuda_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(cuda_lib.YavoError):
        cuda_lib.Context()


def test_host_side_scalars(cuda_lib, oracle):
    import numpy as np
    assert np.array_equal(cuda_lib.ring_points(25, 25), oracle.ring(25, 25))
    d = np.array([5, 9, 10, 19, 20, 40], np.int32)
    assert np.array_equal(cuda_lib.remove_outliers(d, 20), oracle.remove_outliers(d, 20))


def test_product_does_not_import_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ya_vo_b200")
    for dp, dn, fn in os.walk(root):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cc", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in txt and "yavo_oracle" not in txt and "libyavo_oracle" not in txt, f


def test_remove_outliers_empty_train_set_keeps_nothing(cuda_lib, oracle):
    """All distances INT_MAX (empty train set): 2 * INT_MAX wraps to -2 in the reference's int arithmetic
    (src/BriefDescriptor.cc:224), so max(-2, threshold) = threshold and nothing is kept — host C ABI, device filter
    and oracle agree."""
    import numpy as np
    d = np.full(5, 2**31 - 1, np.int32)
    assert not cuda_lib.remove_outliers(d, 20).any()
    assert not oracle.remove_outliers(d, 20).any()
    d = np.array([2**31 - 1, 7, 2**31 - 1, 13, 14], np.int32)
    assert np.array_equal(cuda_lib.remove_outliers(d, 20), oracle.remove_outliers(d, 20))
    assert list(oracle.remove_outliers(d, 20)) == [False, True, False, True, True]


def test_pinned_views_keep_their_block_alive(cuda_lib):
    """Slices handed to callbacks must keep the page-locked block alive after the parent array is dropped."""
    import ctypes as C
    import gc
    import numpy as np
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    freed = []
    ptr = libc.malloc(4096)
    a = cuda_lib._owned_view(ptr, 4096, lambda p: (freed.append(p), libc.free(p))).view(np.int32).reshape(32, 32)
    a[...] = 7
    row = a[3, :5]
    del a
    gc.collect()
    assert freed == [] and int(row.sum()) == 35
    del row
    gc.collect()
    assert freed == [ptr]
