#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_host_gpu.py::test_reference_generator_thread_on_the_b200_evaluator 2>&1 | tail -15
