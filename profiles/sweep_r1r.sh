#!/bin/bash
for epb in 128 32; do echo "DC_EPB=$epb"; DC_EPB=$epb python profiles/chunk_overlap.py exp02_vFinal 65536 1,2,4,8 2>&1 | grep -v Warn; done
