"""Kernels written after the round-1 GPU budget was spent — parallel boundary-search decoders (decode_mixed_kernel,
decode_gorilla_kernel), the warp-parallel Gorilla pre-pass, the parallel V5 run-table reader — selected with
CLDN_B200_UNMEASURED=1 (cldn_kernels.h). They are bit-exact, memcheck- and racecheck-clean under tests/cusim; this file
gives them their hardware run. It sorts LAST on purpose: the hardware-verified defaults are exercised by every other
file first, so a surprise here cannot hide their results behind `pytest -x`.
Every test re-runs a parity test of test_gpu_parity.py / test_gpu_ros.py with the switch on."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import test_gpu_parity as P  # noqa: E402
import test_gpu_ros as R  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _unmeasured_kernels(monkeypatch):
    monkeypatch.setenv("CLDN_B200_UNMEASURED", "1")
    monkeypatch.delenv("CLDN_B200_MIXED_DECODE", raising=False)


def test_golden_vectors(golden, oracle):            # v2 goldens: XOR / Gorilla coders of the reference, wire versions 3 / 4 / 5
    P.test_golden_vectors(golden, oracle)


@pytest.mark.parametrize("version", [5, 4, 3])
@pytest.mark.parametrize("lossless", [True, False])
def test_lossless_float_fields(oracle, version, lossless):   # warp-parallel Gorilla pre-pass + decode_gorilla_kernel / decode_mixed_kernel
    P.test_lossless_float_fields(oracle, version, lossless)


@pytest.mark.parametrize("mode", ["par", "chase"])
@pytest.mark.parametrize("version", [5, 4])
def test_raw_fields_in_the_stream(oracle, monkeypatch, mode, version):
    P.test_raw_fields_in_the_stream(oracle, monkeypatch, mode, version)


def test_gorilla_field_positions(oracle, monkeypatch):
    P.test_gorilla_field_positions(oracle, monkeypatch, "par")


def test_v5_long_run_tables(oracle):                 # parallel Rle / DeltaRle run-table parse
    P.test_v5_long_run_tables(oracle)


@pytest.mark.parametrize("n", [63, 4097, 32775, 100_003])
def test_c3_padded_mixed_sizes(oracle, n):
    P.test_c3_padded_mixed_sizes(oracle, n)


def test_v5_all_modes_and_types(oracle):
    P.test_v5_all_modes_and_types(oracle)


def test_v5_section_decode_errors():
    P.test_v5_section_decode_errors()


def test_skip_store(oracle, monkeypatch):
    P.test_decode_but_skip_store(oracle, monkeypatch, "seq")
    P.test_skip_store_in_v5_sections_is_honoured_here()


def test_c4_velodyne_mixed_layout(oracle):
    P.test_c4_velodyne_mixed_layout(oracle)


def test_corrupted_blobs_decode_like_the_reference(ref):
    P.test_corrupted_blobs_decode_like_the_reference(ref)


def test_dds_paths(ref, golden_ros):                 # the reference's sample layout (FLOAT64 timestamp -> Gorilla) through the envelope
    R.test_dds_roundtrip_like_the_reference_test(golden_ros)
    R.test_compress_message_matches_reference(ref, False)
    R.test_decompress_message_matches_reference(ref)
    R.test_golden_messages(golden_ros)


def test_reference_sample_files(oracle):             # only where the reference tree exists (tests/cusim)
    P.test_reference_sample_files(oracle)
