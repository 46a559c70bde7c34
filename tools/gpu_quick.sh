#!/bin/bash
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/kbench.py 2>&1 | tail -1
