#!/bin/bash
for k in 1 2 4 8; do echo "DC_SUB_BATCHES=$k $(DC_SUB_BATCHES=$k python profiles/quick_time.py 2>&1 | tr '\n' ' ')"; done
