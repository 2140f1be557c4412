#!/bin/bash
mkdir -p gpurun_out
( python scripts/net_time.py; DFVO_PDL=0 python scripts/net_time.py; DFVO_GRAPHS=0 python scripts/net_time.py; DFVO_GRAPHS=0 DFVO_PDL=0 python scripts/net_time.py ) > gpurun_out/net_time.log 2>&1
cat gpurun_out/net_time.log | grep forward
