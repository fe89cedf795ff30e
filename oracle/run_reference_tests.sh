#!/bin/bash
# The reference's OWN test files run against this package (host side; see reference_tests_shim.py).
# Build container only; reads /root/reference/tests in place and runs from a scratch directory (two of the tests
# save models to the working directory), writes nothing under /root/reference.
#   bash oracle/run_reference_tests.sh [pytest args]
here="$(cd "$(dirname "$0")" && pwd)"
ref=/root/reference/tests
[ -d "$ref" ] || exit 1
cd "$(mktemp -d)" || exit 1
files=""
for f in test_window.py test_window_list.py test_parameter.py test_image.py test_image_header.py test_image_list.py \
         test_utils.py test_model.py test_group_models.py test_psfmodel.py test_fit.py; do files="$files $ref/$f"; done
PYTHONDONTWRITEBYTECODE=1 PYTHONPATH="$here:$PYTHONPATH" python -m pytest -p reference_tests_shim -p no:cacheprovider -q -rA "$@" $files
