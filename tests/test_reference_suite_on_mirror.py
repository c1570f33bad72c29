"""Runs the REFERENCE's own unit tests (framework/register_test.py, config_test.py, module_test.py -
the only tests the reference ships, 21 cases) against this repo's mirror of the framework: the
test files are loaded from where they lie under /root/reference (nothing is copied) with the
``framework`` package name resolved to ``rag_arc_b200.framework``.  Skipped where the reference
tree does not exist (the GPU box)."""
import importlib
import importlib.util
import os
import sys
import unittest

import pytest

REF = os.environ.get("RAGARC_REFERENCE", "/root/reference")
FILES = ["register_test.py", "config_test.py", "module_test.py"]


@pytest.mark.parametrize("fname", FILES)
def test_reference_framework_tests_pass_on_the_mirror(fname):
    path = os.path.join(REF, "framework", fname)
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    saved = {k: v for k, v in sys.modules.items() if k == "framework" or k.startswith("framework.")}
    try:
        mirror = importlib.import_module("rag_arc_b200.framework")
        sys.modules["framework"] = mirror
        for sub in ("register", "module", "config", "singleton_decorator"):
            sys.modules[f"framework.{sub}"] = importlib.import_module(f"rag_arc_b200.framework.{sub}")
        spec = importlib.util.spec_from_file_location(f"_ref_{fname[:-3]}", path)
        mod = importlib.util.module_from_spec(spec)
        sys.dont_write_bytecode = True
        spec.loader.exec_module(mod)
        suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
        assert suite.countTestCases() > 0
        result = unittest.TextTestRunner(stream=open(os.devnull, "w"), verbosity=0).run(suite)
        problems = [f"{t}: {tb.splitlines()[-1]}" for t, tb in result.failures + result.errors]
        assert result.wasSuccessful(), f"{fname}: {problems}"
    finally:
        for k in [k for k in sys.modules if k == "framework" or k.startswith("framework.")]:
            del sys.modules[k]
        sys.modules.update(saved)
