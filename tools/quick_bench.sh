#!/bin/bash
# usage: tools/quick_bench.sh "<nvcc extra flags>" : rebuild with flags, run a short bench, print value
cd "$(dirname "$0")/.."
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $1 -o mpp_cnn_rs_object_detection_b200/libmpp_b200.so mpp_cnn_rs_object_detection_b200/csrc/mpp_b200.cu || exit 1
touch mpp_cnn_rs_object_detection_b200/libmpp_b200.so
