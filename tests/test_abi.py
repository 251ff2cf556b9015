"""The drop-in boundary, checked without a GPU: the shared library loads, exports every symbol the C-ABI header
declares and the reference's C++-mangled entry points; struct layouts equal the reference build's; the host-side
scanner session logic agrees with the oracle's restatement of scan.cpp."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from util import ROOT, deck_frames

LIB = os.path.join(ROOT, "card.io-dmz_b200", "libb200dmz.so")

# mangled names of the reference's own build (nm -D oracle/_ref/libdmz_ref.so)
REFERENCE_SYMBOLS = [
    "_Z18dmz_context_createv", "_Z19dmz_context_destroyP11dmz_context", "_Z19dmz_found_all_edges9dmz_edges",
    "_Z16dmz_detect_edgesP9_IplImageS0_S0_hP9dmz_edgesP17dmz_corner_points",
    "_Z18dmz_transform_cardP11dmz_contextP9_IplImage17dmz_corner_pointshbPS2_",
    "_Z18scanner_initializeP12ScannerState", "_Z13scanner_resetP12ScannerState",
    "_Z17scanner_add_frameP12ScannerStateP9_IplImageP15FrameScanResult",
    "_Z29scanner_add_frame_with_expiryP12ScannerStateP9_IplImagebP15FrameScanResult",
    "_Z14scanner_resultP12ScannerStateP13ScannerResult", "_Z15scanner_destroyP12ScannerState",
    "_Z25dmz_deinterleave_uint8_c2P9_IplImagePS0_S1_", "_Z19dmz_best_expiry_segP9_IplImagetPP18CythonGroupedRectsPt", "_Z15dmz_focus_scoreP9_IplImageb", "_Z20dmz_brightness_scoreP9_IplImageb",
    "_Z14dmz_has_opencvv", "_Z16dmz_YCbCr_to_RGBP9_IplImageS0_S0_PS0_", "_Z26dmz_deinterleave_RGBA_to_RPhS_i",
    "_Z18dmz_scharr3_dx_absP9_IplImageS0_", "_Z18dmz_scharr3_dy_absP9_IplImageS0_", "_Z16dmz_sobel3_dx_dyP9_IplImageS0_",
]


def exported():
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB], text=True)
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_library_loads_without_gpu():
    assert os.path.exists(LIB), "run __graft_entry__.build() first"
    C.CDLL(LIB)


def test_exports_every_declared_c_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200_dmz.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 20
    missing = declared - exported()
    assert not missing, missing


def test_header_is_plain_c99_and_links(tmp_path):
    """include/b200_dmz.h compiled by a strict C99 compiler, the caller linked with the library and run (host-side calls only)."""
    exe = str(tmp_path / "c_abi_main")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_main.c"), "-o", exe, "-L" + os.path.dirname(LIB), "-lb200dmz",
                           "-Wl,-rpath," + os.path.dirname(LIB)])
    out = subprocess.check_output([exe], text=True).split()
    assert out[:3] == ["0", "0", "808"]  # not complete, no digits, sizeof(b200_frame_record)


def test_exports_reference_cxx_entry_points():
    missing = set(REFERENCE_SYMBOLS) - exported()
    assert not missing, missing


def test_reference_symbol_names_are_current(ref):
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "oracle", "_ref", "libdmz_ref.so")], text=True)
    have = {line.split()[-1] for line in out.splitlines() if line.strip()}
    assert set(REFERENCE_SYMBOLS) <= have


def test_struct_layouts(pkg, ref):
    # sizes asserted at compile time in include/dmz_b200_compat.h against these reference-build numbers
    sizes = [ref.lib.ref_sizeof(i) for i in range(10)]
    assert sizes == [28, 48, 640, 816, 240, 2832, 48, 32, 144, 520]
    assert C.sizeof(pkg.VSeg) == sizes[0] and C.sizeof(pkg.HSeg) == sizes[1]
    assert C.sizeof(pkg.Edges) == sizes[6] and C.sizeof(pkg.CornerPoints) == sizes[7]


def test_member_layouts_match_reference_headers(tmp_path):
    """offsetof / sizeof of EVERY member that crosses the boundary: include/dmz_b200_compat.h against the listing the same
    probe (tests/layout_probe.cpp) printed when compiled against the reference's own unmodified dmz.h + scan/scan.h
    (tests/golden/ref_layout.txt, written by `make -C oracle layout`; regenerated here when /root/reference exists)."""
    golden = os.path.join(ROOT, "tests", "golden", "ref_layout.txt")
    want = open(golden).read()
    ref_dir = os.environ.get("DMZ_REFERENCE", "/root/reference")
    if os.path.isdir(ref_dir):
        exe = str(tmp_path / "probe_ref")
        subprocess.check_call(["g++", "-std=gnu++03", "-w", "-DCYTHON_DMZ=1", "-DPROBE_REFERENCE_HEADERS",
                               "-I" + os.path.join(ROOT, "oracle", "stub_include"), "-I" + ref_dir,
                               os.path.join(ROOT, "tests", "layout_probe.cpp"), "-o", exe])
        assert subprocess.check_output([exe], text=True) == want, "tests/golden/ref_layout.txt is stale: make -C oracle layout"
    exe = str(tmp_path / "probe_b200")
    subprocess.check_call(["g++", "-std=c++14", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "layout_probe.cpp"), "-o", exe])
    got = subprocess.check_output([exe], text=True)
    assert len(want.splitlines()) > 120
    assert got == want


def test_tensor_core_operand_tables(tmp_path):
    """The host-built operand tables of the tensor-core kernels (b200_tables.cpp) against the float weights: vseg W1 as four
    base-128 digits (exact to 2^-27 of the unit's largest weight, padding zero), the (s, d0) table against the reference's
    three-rounding normalisation for every (min, max, value), CNN conv taps as three digits placed at the right
    (pool position, window byte) with 255 * sum |Q| < 2^31, hidden weights as fp16 hi + lo."""
    exe = str(tmp_path / "tables_main")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"),
                           "-I" + os.path.join(ROOT, "card.io-dmz_b200", "csrc"), "-I/usr/local/cuda/include",
                           os.path.join(ROOT, "tests", "tables_main.cpp"), os.path.join(ROOT, "card.io-dmz_b200", "csrc", "b200_tables.cpp"), "-o", exe])
    out = subprocess.run([exe, os.path.join(ROOT, "card.io-dmz_b200", "weights")], capture_output=True, text=True)
    vals = dict(line.split() for line in out.stdout.splitlines())
    assert out.returncode == 0 and vals["failed_checks"] == "0", out.stdout
    assert float(vals["vseg_weight_rel_err"]) < 8e-9 and float(vals["cnn_conv_weight_rel_err"]) < 1e-6


def test_no_context_without_cuda(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.B200Error, match="no CPU fallback"):
        pkg.Dmz()


def test_scanner_session_matches_oracle(pkg, oracle):
    """b200_scanner_* (host logic, scan.cpp:41-194) fed with oracle scan records == the oracle's own session."""
    frames = deck_frames(16, 8)
    recs, cards = oracle.process_frames(frames, want_cards=True)
    so = oracle.scanner_new()
    sb = pkg.Scanner()
    for k in range(8):
        oracle.scanner_add_frame(so, cards[k])
        sb.add_scan(recs[k])
        a15o, a16o, cnto = oracle.scanner_peek(so)
        a15b, a16b, cntb = sb.peek()
        assert np.array_equal(cnto, cntb)
        assert np.array_equal(a16o.view(np.uint32), a16b.view(np.uint32)) and np.array_equal(a15o.view(np.uint32), a15b.view(np.uint32))
        do, go = oracle.scanner_result(so)
        db, gb = sb.result()
        assert do == db and go.tolist() == gb.tolist()
    assert db, "the deck session should complete"
    oracle.scanner_free(so)
    sb.close()
