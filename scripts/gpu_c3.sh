#!/bin/bash
# One gpurun call: the C3 registration test and the full-size C3 script (grad-NCC and patch grad-NCC).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_registration.py -m gpu -x -q > gpurun_out/pytest_regi.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_regi.log
tail -15 gpurun_out/pytest_regi.log
timeout 600 python scripts/register_c3.py > gpurun_out/register_c3.jsonl 2> gpurun_out/register_c3.err; tail -4 gpurun_out/register_c3.jsonl; tail -3 gpurun_out/register_c3.err
timeout 600 python scripts/register_c3.py --metric patch-grad-ncc > gpurun_out/register_c3_pgncc.jsonl 2> gpurun_out/register_c3_pgncc.err; tail -4 gpurun_out/register_c3_pgncc.jsonl; tail -3 gpurun_out/register_c3_pgncc.err
