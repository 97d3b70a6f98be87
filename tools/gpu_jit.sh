#!/bin/bash
timeout 300 python tools/fused_time.py cfg3 65536 2 14 2>&1 | tail -2
