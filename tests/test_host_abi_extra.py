"""CPU-side checks of the host-only ABI additions of round 2."""
import ctypes as C


def test_exact_domain_query(pkg):
    """sw_params_in_exact_domain: match + gap_open <= 0 is the RTL's schedule-independent domain
    (SW_ProcessingElement_v1.0.v:120 vs :131-141)."""
    assert pkg.params_in_exact_domain() == 1                              # 5/-4/-12/-4
    assert pkg.params_in_exact_domain(5, -4, -8, -4) == 1                 # the swalign vectors' set
    assert pkg.params_in_exact_domain(5, -4, -5, -1) == 1                 # boundary: match + gap_open == 0
    assert pkg.params_in_exact_domain(5, -4, -2, -1) == 0                 # accepted, RTL schedule-dependent
    assert pkg.params_in_exact_domain(5, -4, -12, 3) == pkg.SW_EINVAL     # rejected by sw_init as well
    assert pkg.load_library().sw_params_in_exact_domain(None) == 1


def test_stats_struct_layout(pkg):
    assert C.sizeof(pkg.SwStats) == 6 * 8


def test_strip_kernel_source_compiles_under_nvrtc(pkg, tmp_path, monkeypatch):
    """The run-time specialisation path (csrc/sw_jit.cu) compiles the embedded strip-kernel source with
    NVRTC.  The compile stage needs no GPU, so it is checked here: without a device the call gets as
    far as loading the cubin and fails THERE, never in the compiler."""
    monkeypatch.setenv("SW_B200_JIT_CACHE", str(tmp_path))
    for variant in ("strip_s16x2_R25x2_G1", "strip_s16x2_R25x2_G1_U4_F31", "strip_s16x2_R16x1_G32"):
        ok, msg = pkg.jit_compile_check(variant, -7, -3)
        if "could not be loaded" in msg:
            import pytest
            pytest.skip("NVRTC is not installed here")
        assert ok == 1 or "loading the specialised cubin failed" in msg, (variant, msg)
        assert "NVRTC compile failed" not in msg


def test_switch_setters_reject_a_null_handle(pkg):
    """The plan / work-order switches validate their handle without touching a device."""
    lib = pkg.load_library()
    assert lib.sw_set_pass_split(None, 0) == pkg.SW_EINVAL
    assert lib.sw_last_pass_parts(None) == pkg.SW_EINVAL
    assert lib.sw_set_launch_plan(None, 2, 0) == pkg.SW_EINVAL
    assert lib.sw_set_wave_mode(None, 1) == pkg.SW_EINVAL


def test_pass_split_arithmetic(pkg):
    """sw_plan_pass_parts: the host arithmetic of the pass split (csrc/sw_api.cu plan_pass_split).  Parts are
    whole profile chunks; the automatic rule splits launches of 1 .. 16 rounds of multi-chunk items into
    about 20 rounds of part-items; shapes as measured in profiles/r02_pass_split_ab.txt."""
    # 200 k x 1 kb subjects x one 10 kb query on R38x2_G1: 132 passes in chunks of 5, 782 chains, 296 blocks
    ok, parts, pp = pkg.plan_pass_parts(132, 5, 782, 296)
    assert ok and pp % 5 == 0 and parts == -(-132 // pp) and parts == 9
    assert 782 * parts / 296 >= 20
    # 1 M subjects: 13.2 rounds -> a few parts only
    ok, parts, pp = pkg.plan_pass_parts(132, 5, 3907, 296)
    assert ok and parts == 3 and pp == 65
    # many rounds already / under-filled GPU / single profile chunk / switched off: not split
    assert pkg.plan_pass_parts(132, 5, 296 * 16, 296) == (False, 1, 132)
    assert pkg.plan_pass_parts(132, 5, 200, 296) == (False, 1, 132)
    assert pkg.plan_pass_parts(3, 7, 39063, 444) == (False, 1, 3)
    assert pkg.plan_pass_parts(132, 5, 782, 296, mode=0) == (False, 1, 132)
    # forced: about n parts, never finer than one chunk per part, every pass covered exactly once
    for npass, cp, want in ((60, 7, 3), (60, 7, 64), (5, 1, 2), (31, 4, 5)):
        ok, parts, pp = pkg.plan_pass_parts(npass, cp, 10, 296, mode=want)
        assert ok and pp % cp == 0 and (parts - 1) * pp < npass <= parts * pp
        assert parts <= -(-npass // cp)
    # the 32-bit work counter: parts x chains must stay below 2^31
    assert pkg.plan_pass_parts(1000, 1, 1 << 30, 296, mode=8)[0] is False
