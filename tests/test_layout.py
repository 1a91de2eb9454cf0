"""Repository rules: the product never touches the oracle, and never reads /root/reference."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_and_cu(path):
    for d, _dirs, files in os.walk(path):
        if "_build" in d or "__pycache__" in d:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                yield os.path.join(d, f)


def test_product_does_not_import_or_link_the_oracle():
    bad = []
    for f in list(_py_and_cu(os.path.join(ROOT, "adcraft_b200"))) + [os.path.join(ROOT, "include", "adcraft_b200.h")]:
        text = open(f).read()
        if re.search(r"^\s*(from|import)\s+oracle\b|oracle[./]|adcraft_oracle|liboracle", text, flags=re.M):
            bad.append(f)
    assert not bad, f"product files reference the oracle: {bad}"


def test_nothing_shipped_reads_the_reference_tree_at_run_time():
    for f in list(_py_and_cu(os.path.join(ROOT, "adcraft_b200"))) + [
            os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]:
        if os.path.exists(f):
            assert "/root/reference" not in open(f).read(), f


def test_required_top_level_layout():
    for p in ("bench.py", "__graft_entry__.py", "DESIGN.md", "INTEGRATION.md", "include/adcraft_b200.h",
              "oracle/adcraft_oracle.c", "oracle/Makefile", "tests/golden/make_golden.py", "profiles"):
        assert os.path.exists(os.path.join(ROOT, p)), p


def test_sampler_tables_of_library_and_oracle_are_the_same_literals():
    """The Exp(1) and normal sampler tables are spec constants: the CUDA library and the C oracle
    each keep a copy (tools/gen_neglog_table.py, tools/gen_znorm_table.py write both), and bit-exact
    free-running parity rests on the copies being identical."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for lib, orc, rows in (("adc_neglog_table.inc", "neglog_table.inc", 128),
                           ("adc_znorm_table.inc", "znorm_table.inc", 32 * 128)):
        a = open(os.path.join(root, "adcraft_b200", "csrc", lib)).read()
        b = open(os.path.join(root, "oracle", orc)).read()
        assert a == b, (lib, orc)
        assert sum(1 for line in a.splitlines() if line.strip().startswith("{")) == rows, lib
