#!/bin/bash
t() { echo "== $1"; env $2 timeout 200 python -m pytest tests/test_llama_gpu.py -m gpu -q -k "dual and d64" --timeout 200 -p no:cacheprovider 2>&1 | grep -E "passed|failed|assert 0\." | head -3 | cut -c1-120; }
t "default" "X=1"
t "fuse=1" "ZL_TEST_FUSE=1"
t "force half kernels" "ZL_W4_FORCE_HALF=1"
t "prompt 32 (one chunk 16/16)" "ZL_TEST_PROMPT=32"
t "prompt 16 (8/8 PACK? no: halves of 8)" "ZL_TEST_PROMPT=16"
t "no pack" "ZL_W4_NO_ONE=1"
