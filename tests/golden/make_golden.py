"""Regenerates the fixtures in this directory.  Run in the build container only (it reads /root/reference):

    python tests/golden/make_golden.py

* employees.csv / departments.csv — the reference's own sample tables (data/employees.csv:1-7, data/departments.csv:1-5),
  the only fixtures the reference ships for this path.
* employees_golden.json — hand-derived expected results on them (SURVEY.md 8c "Golden vectors"): the reference has no
  test that pins them and cannot be run here (no cargo/rustc), so the values follow from reading the Rust:
  schema inference makes dept_id Utf8 (row 6 holds the text NULL), age > 25 keeps Bob, Charlie, Diana, Eve, Frank.
* siphash13_golden.json — SipHash-1-3 with the all-zero key (Rust `DefaultHasher::new()`), produced by CPython's own
  siphash13 (`hash(bytes)` under PYTHONHASHSEED=0): an implementation independent of the oracle's.
"""
import json
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/data")


def main():
    for name in ("employees.csv", "departments.csv"):
        shutil.copyfile(REF / name, HERE / name)
    golden = {
        "tier1_scalar": {
            "sql": "SELECT COUNT(id), SUM(salary), AVG(salary), MIN(age), MAX(age) FROM employees WHERE age > 25",
            "columns": ["col_0", "col_1", "col_2", "col_3", "col_4"],
            "types": ["int64", "int64", "double", "int64", "int64"],
            "rows": [[5, 428000, 85600.0, 28, 35]],
        },
        "tier2_group_by": {
            "sql": "SELECT dept_id, COUNT(id), SUM(salary), AVG(salary) FROM employees WHERE age > 25 GROUP BY dept_id",
            "columns": ["dept_id", "col_0", "col_1", "col_2"],
            "types": ["string", "int64", "int64", "double"],
            "rows": [["101", 1, 95000, 95000.0], ["102", 2, 175000, 87500.0], ["103", 1, 80000, 80000.0], ["NULL", 1, 78000, 78000.0]],
            "reference_asis_rows": [],
        },
        "filter_age_gt_25": {"ids": [2, 3, 4, 5, 6]},
    }
    (HERE / "employees_golden.json").write_text(json.dumps(golden, indent=1) + "\n")

    inputs = [b"a", b"abc", b"hello world", bytes(range(7)), bytes(range(8)), bytes(range(9)), bytes(range(15)),
              bytes(range(16)), bytes(range(64)), (1).to_bytes(8, "little"), (-1).to_bytes(8, "little", signed=True),
              (42).to_bytes(4, "little"), b"101\xff", b"NULL\xff", (1).to_bytes(8, "little") + b"x\xff"]
    code = ("import sys,json;assert sys.hash_info.algorithm=='siphash13';"
            "ins=[bytes.fromhex(h) for h in json.loads(sys.argv[1])];"
            "print(json.dumps([hash(b) % (1<<64) for b in ins]))")
    out = subprocess.run([sys.executable, "-c", code, json.dumps([b.hex() for b in inputs])], capture_output=True, text=True,
                         env=dict(os.environ, PYTHONHASHSEED="0"), check=True).stdout
    hashes = json.loads(out)
    # CPython special-cases hash(b'') == 0 (not SipHash) and maps -1 to -2; the inputs above avoid both
    (HERE / "siphash13_golden.json").write_text(json.dumps({"key": [0, 0], "vectors": [[b.hex(), h] for b, h in zip(inputs, hashes)]}, indent=1) + "\n")


if __name__ == "__main__":
    main()
