#!/bin/bash
for cfg in "128 32" "64 32" "96 24"; do set -- $cfg; echo "DC_EPB=$1 DC_EPW=$2 $(DC_EPB=$1 DC_EPW=$2 python profiles/quick_time.py 2>&1 | head -1)"; done
