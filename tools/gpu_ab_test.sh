#!/bin/bash
# usage: tools/gpu_ab_test.sh <variant> "<pytest -k expr>" "<bench args>"  — parity tests with a variant library, then A/B against the in-tree one
v=$1; k=$2; args=$3
AMDP_B200_LIB=$PWD/build/variants/libamdp_$v.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_policy.py -m gpu -x -q -k "$k" 2>&1 | tail -3
tools/gpu_var.sh "$args" base $v
