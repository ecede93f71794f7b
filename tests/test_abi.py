"""The drop-in boundary: libsiril_stack_cuda.so loads, exports every symbol that
include/siril_stack_cuda.h declares, keeps the reference's struct stacking_args layout and
enums, and fails loudly (no CPU fallback) when no CUDA device is present.  No compute here."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "siril_stack_cuda.h")
REFERENCE = "/root/reference/src"


def test_library_exports_every_declared_symbol(pkg):
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(ssc_[a-z0-9_]+)\s*\(", text))
    declared -= {"ssc_callbacks"}
    assert len(declared) >= 20
    assert declared == set(pkg.EXPORTS)
    for name in declared:
        assert getattr(pkg.lib(), name) is not None, name


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 with nothing but libc headers."""
    src = tmp_path / "t.c"
    src.write_text('#include "siril_stack_cuda.h"\nint main(void){struct stacking_args a; ssc_band b; '
                   'ssc_result r; (void)a;(void)b;(void)r; return sizeof(struct stacking_args)==224?0:1;}\n')
    exe = tmp_path / "t"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    assert subprocess.call([str(exe)]) == 0


def test_stacking_args_layout_matches_ctypes(pkg):
    # measured x86-64 LP64 layout of the reference struct
    expect = dict(method=0, seq=8, filtering_criterion=16, filtering_parameter=24, nb_images_to_stack=32,
                  image_indices=40, description=48, output_filename=152, output_overwrite=160, t_start=168,
                  retval=184, max_number_of_rows=188, sig=192, type_of_rejection=208, normalize=212,
                  force_norm=216)
    assert C.sizeof(pkg.StackingArgs) == 224
    for k, v in expect.items():
        assert getattr(pkg.StackingArgs, k).offset == v, k


@pytest.mark.skipif(not os.path.exists(REFERENCE), reason="reference sources not on this machine")
def test_layout_and_enums_match_reference_headers(tmp_path):
    """Compile the same probe against the reference's stacking.h and against ours."""
    probe = r'''
#include <stdio.h>
#include <stddef.h>
%s
int main(void) {
    printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu\n", sizeof(struct stacking_args),
        offsetof(struct stacking_args, seq), offsetof(struct stacking_args, image_indices),
        offsetof(struct stacking_args, description), offsetof(struct stacking_args, t_start),
        offsetof(struct stacking_args, max_number_of_rows), offsetof(struct stacking_args, sig),
        offsetof(struct stacking_args, type_of_rejection), offsetof(struct stacking_args, force_norm));
    printf("%%d %%d %%d %%d %%d %%d | %%d %%d %%d %%d %%d | %%zu\n", NO_REJEC, PERCENTILE, SIGMA, SIGMEDIAN,
        WINSORIZED, LINEARFIT, NO_NORM, ADDITIVE, MULTIPLICATIVE, ADDITIVE_SCALING, MULTIPLICATIVE_SCALING,
        sizeof(norm_coeff));
    return 0;
}
'''
    outs = []
    stubs = os.path.join(ROOT, "oracle", "ref_shim", "stubs")
    for tag, inc, flags in (
            ("ref", '#include "core/siril.h"\n#include "stacking/stacking.h"', ["-I", stubs, "-I", REFERENCE, "-w"]),
            ("ours", '#include "siril_stack_cuda.h"', ["-I", os.path.join(ROOT, "include")])):
        src = tmp_path / (tag + ".c")
        src.write_text(probe % inc)
        exe = tmp_path / tag
        subprocess.check_call(["/usr/bin/gcc", str(src), "-o", str(exe)] + flags)
        outs.append(subprocess.check_output([str(exe)]).decode())
    assert outs[0] == outs[1]


def test_header_cites_reference_lines():
    text = open(HEADER).read()
    for cite in ("stacking.h:38-56", "stacking.c:54-56", "stacking.c:196", "stacking.c:125", "sequence.c:519"):
        assert cite in text, cite


def test_no_cpu_fallback(pkg):
    """Without a device the operators fail with SSC_ERR_CUDA -- they never compute on the host."""
    if pkg.device_count() > 0:
        pytest.skip("a CUDA device is present")
    frames = np.zeros((3, 1, 8, 8), np.uint16)
    seq = pkg.Sequence(frames)
    for fn in (pkg.stack_summing, pkg.stack_median, pkg.stack_mean_with_rejection, pkg.stack_addmax,
               pkg.stack_addmin):
        with pytest.raises(pkg.StackError) as e:
            fn(seq)
        assert e.value.code == pkg.ERR_CUDA


def test_argument_errors_keep_reference_codes(pkg):
    frames = np.zeros((3, 1, 8, 8), np.uint16)
    seq = pkg.Sequence(frames)
    args = pkg.make_args(seq, image_indices=[0])           # a single frame: the reference returns -1
    assert pkg.lib().ssc_stack_median(C.byref(args)) == -1
    assert pkg.lib().ssc_stack_mean_with_rejection(C.byref(args)) == -1
    assert pkg.lib().ssc_stack_summing(C.byref(args)) == -1
    assert pkg.lib().ssc_stack_addmax(C.byref(args)) == -1
    assert pkg.lib().ssc_stack_addmin(C.byref(args)) == -1
    assert args.retval == -1
    args = pkg.make_args(seq)
    args.seq = 0xdead0                                      # never bound
    assert pkg.lib().ssc_stack_median(C.byref(args)) == pkg.ERR_UNBOUND
    r = pkg.Result()
    assert pkg.lib().ssc_take_result(C.byref(r)) == pkg.ERR_NO_RESULT


def test_product_never_touches_the_oracle():
    """The product path must not import, link or open anything under oracle/."""
    pkgdir = os.path.join(ROOT, "siril-0.9_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "libsiril_oracle" not in text and "libsiril_ref" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
    needed = subprocess.check_output(["ldd", os.path.join(pkgdir, "libsiril_stack_cuda.so")]).decode()
    assert "oracle" not in needed and "siril_ref" not in needed
