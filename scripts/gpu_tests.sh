# GPU parity tests only (optionally -k expr)
timeout 1500 python -m pytest tests -x -q -m gpu ${1:+-k "$1"} 2>&1 | tail -${TAILN:-30}
