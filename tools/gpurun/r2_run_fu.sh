#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 300 python tools/step_time.py --steps 20 --shipped --profile 2>&1 | tail -2
NINT_FUSE_STEPS=0 timeout 300 python tools/step_time.py --steps 20 --shipped --profile 2>&1 | tail -2
