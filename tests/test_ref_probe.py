"""CPU tests of the reference-package probe (tests/ref_probe.py): it must find a real build when one is dropped
into baseline/_ref or named by FSGS_REF_RASTERIZER, and must never mistake this repository's own drop-in packages
for the reference."""
import os
import sys

sys.path.insert(0, os.path.dirname(__file__))
import ref_probe  # noqa: E402


def test_our_own_packages_are_never_reported_as_the_reference():
    found = ref_probe.find_reference_rasterizer()
    if found is not None:       # a real build is installed on this machine: fine, but it must not be ours
        assert not any(found.startswith(o) for o in ref_probe.OURS)
    for o in ref_probe.OURS:
        pkg = os.path.join(o, ref_probe.NAME)
        assert os.path.isfile(os.path.join(pkg, "__init__.py")), pkg
        assert not ref_probe._has_compiled_C(pkg)


def test_probe_finds_a_package_with_a_compiled_extension(tmp_path, monkeypatch):
    pkg = tmp_path / "somewhere" / ref_probe.NAME
    pkg.mkdir(parents=True)
    (pkg / "__init__.py").write_text("from . import _C\n")
    monkeypatch.setenv("FSGS_REF_RASTERIZER", str(tmp_path / "somewhere"))
    assert ref_probe.find_reference_rasterizer() != str(pkg)          # no compiled _C yet: not a candidate
    (pkg / "_C.cpython-312-x86_64-linux-gnu.so").write_bytes(b"not really an ELF file")
    assert ref_probe.find_reference_rasterizer() == str(pkg)
    # a build that does not load on this machine is reported and skipped, it never breaks the suite
    assert ref_probe.load_reference_rasterizer(str(pkg)) is None
    import diff_gaussian_rasterization as ours                          # and ours is still what the name resolves to
    assert os.path.abspath(ours.__file__).startswith(ref_probe.OURS[0])
