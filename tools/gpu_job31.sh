#!/bin/bash
timeout 600 python tools/k4_determinism.py 2>&1 | tail -30
