#!/bin/bash
# Install the UNMODIFIED reference into baseline/_ref/ (git-ignored; it travels to the GPU box with the snapshot).
# Run once in the build container: the reference tree is read-only, so it is installed from a copy under /tmp.
#     bash oracle/vendor_reference.sh
# bench.py --impl reference (and the cpu_baseline of the default run) then time the reference itself through
# oracle/ref_shim.py + oracle/ref_runner.py and say `"kind": "reference"`; without baseline/_ref/ they fall back to the
# bit-checked port (oracle/ilqr_numpy.py) and say `"kind": "port"`.
set -e
cd "$(dirname "$0")/.."
SRC=${CURIOUSRL_REFERENCE_SRC:-/root/reference}
rm -rf /tmp/curiousrl_ref_src baseline/_ref
cp -r "$SRC" /tmp/curiousrl_ref_src
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
    --target baseline/_ref /tmp/curiousrl_ref_src
ls baseline/_ref
