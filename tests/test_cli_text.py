"""CPU: readme.txt and the parameter block of log.txt equal what the reference's unmodified
scripts/run_midas.py writes (tests/golden/cli/, made by tests/golden/make_golden_cli.py)."""
import io
import json
import os
import sys

import pytest

from midas_b200.run import cli_text

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "cli")
with open(os.path.join(GOLD, "cases.json")) as fh:
    SPEC = json.load(fh)


@pytest.mark.parametrize("name", sorted(SPEC["cases"]))
def test_log_parameter_block_and_readme(name, tmp_path, monkeypatch, capsys):
    program, a = SPEC["cases"][name]
    os.makedirs(tmp_path / program)
    args = dict(a, outdir=str(tmp_path), db="/db/midas_db_v1.2", log=io.StringIO())
    monkeypatch.setattr(sys, "argv", list(SPEC["argv"]))
    cli_text.print_arguments(program, args)
    want = open(os.path.join(GOLD, name + ".params.txt")).read()
    assert args["log"].getvalue().replace(str(tmp_path), "OUT") == want
    assert capsys.readouterr().out.replace(str(tmp_path), "OUT") == want          # the block also goes to stdout
    cli_text.write_readme(program, args)
    assert open(tmp_path / program / "readme.txt").read() == open(os.path.join(GOLD, name + ".readme.txt")).read()
