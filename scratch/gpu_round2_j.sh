#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scratch/ab_libs.py base=sea_ice_drift_b200/libsid_b200.so ns2=scratch/ab/libsid_ns2.so ns3=scratch/ab/libsid_ns3.so cfg2 cfg3 > gpurun_out/r2j_ab.txt 2>&1; cat gpurun_out/r2j_ab.txt
