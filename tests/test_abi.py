"""The C-ABI library loads without a GPU and exports every symbol include/dcue_b200.h declares
(no compute calls here)."""
import ctypes
import importlib
import os

pkg = importlib.import_module("amplifai-deepcontentrecommenders_b200")
L = pkg._lib


def test_library_exports_every_declared_symbol():
    protos = L.parse_header()
    assert len(protos) >= 30
    handle = ctypes.CDLL(L.LIB_PATH)
    for name in protos:
        assert getattr(handle, name) is not None, name      # AttributeError if missing
    assert L.lib().dcue_version() == 100
    assert L.lib().dcue_launch_count() == 0                  # nothing has run on this CPU-only box


def test_header_is_the_single_source_of_prototypes():
    protos = L.parse_header()
    res, args, names = protos["dcue_score_hinge_fwdbwd"]
    assert res is ctypes.c_int and len(args) == 13 and names[-1] == "stream"
    assert protos["dcue_conv_ws_bytes"][0] is ctypes.c_size_t
    # argument validation happens before any CUDA call: a NULL pointer is rejected with a message
    rc = L.lib().dcue_score_fwd(None, None, 1, 1, 100, 1e-8, None, None)
    assert rc == -1 and b"bad argument" in L.lib().dcue_last_error()


def test_alias_module_and_public_api():
    import dcue_b200
    assert dcue_b200.DCUE is pkg.DCUE and dcue_b200.DCUENet is pkg.DCUENet
    from dcue_b200.parallel import DataParallelDCUE, shard_slice  # noqa: F401
    assert os.path.basename(os.path.dirname(dcue_b200.__file__)) == "amplifai-deepcontentrecommenders_b200"
