#!/bin/bash
# The reference's OWN test files run against this package (host side; see reference_tests_shim.py).
# Build container only; reads /root/reference/tests in place, writes nothing there.
#   bash oracle/run_reference_tests.sh [pytest args]
here="$(cd "$(dirname "$0")" && pwd)"
cd /root/reference/tests || exit 1
files="test_window.py test_window_list.py test_parameter.py test_image.py test_image_header.py test_image_list.py test_utils.py test_model.py test_group_models.py test_psfmodel.py test_fit.py"
PYTHONDONTWRITEBYTECODE=1 PYTHONPATH="$here:$PYTHONPATH" python -m pytest -p reference_tests_shim -p no:cacheprovider -q -rA "$@" $files
