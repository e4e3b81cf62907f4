#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -x -q -k "k5 or multih_class_surface" 2>&1 | tail -6
timeout 60 python tools/nb_time.py 2>&1 | tail -6
