#!/bin/bash
cp /tmp/x 2>/dev/null
python tools/mt_time.py
timeout 1500 python -m pytest tests -x -q -m gpu -k "mt or deterministic or reference or golden or tree or compat" 2>&1 | tail -3
