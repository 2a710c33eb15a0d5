#!/bin/bash
timeout 600 python tools/tc_epi_bench.py 2>&1 | tail -30
