#!/bin/bash
mkdir -p gpurun_out
cp zig_vulkan_b200/libvrt.so /tmp/libvrt_orig.so; cp build/ab/libvrt_stats.so zig_vulkan_b200/libvrt.so
timeout -k 5 200 python tools/gpu_tilestats.py C3 2>&1 | tee gpurun_out/tilestats_C3.log
cp /tmp/libvrt_orig.so zig_vulkan_b200/libvrt.so
