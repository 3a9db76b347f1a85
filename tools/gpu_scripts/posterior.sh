#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/posterior_agreement.py > gpurun_out/posterior_agreement.log 2>&1; echo rc=$?; grep -v "Average acceptance\|MCMC scale\|^$" gpurun_out/posterior_agreement.log | tail -8 | cut -c1-330
