#!/bin/bash
# Install the UNMODIFIED reference into baseline/_ref (git-ignored) for `bench.py --impl reference`.
# The reference ships neither setup.py nor pyproject.toml, so `pip install /root/reference` fails with
# "Directory is not installable"; we install from a /tmp copy that only ADDS a packaging shim (setup.py
# listing the reference's own files).  No reference source file is changed; the script verifies that.
set -euo pipefail
cd "$(dirname "$0")/.."
rm -rf /tmp/refsrc baseline/_ref
cp -r /root/reference /tmp/refsrc
cat > /tmp/refsrc/setup.py <<'PY'
from setuptools import setup
setup(name="dlb-reference", version="0.0.0",
      py_modules=["dbs", "dataloader", "parser", "dbs_logging", "utils", "prepare_data"],
      packages=["Net"],
      data_files=[("rnn_data/wikitext-2", ["rnn_data/wikitext-2/train.txt", "rnn_data/wikitext-2/valid.txt",
                                           "rnn_data/wikitext-2/test.txt"])])
PY
touch /tmp/refsrc/Net/__init__.py
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target baseline/_ref /tmp/refsrc
rm -f baseline/_ref/Net/__init__.py          # the reference uses Net as a namespace package
for f in dbs dataloader parser dbs_logging utils; do cmp /root/reference/$f.py baseline/_ref/$f.py; done
diff -r -x __pycache__ /root/reference/Net baseline/_ref/Net
echo "reference installed unmodified into baseline/_ref"
