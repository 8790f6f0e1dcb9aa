#!/bin/bash
# host topology of the GPU box: NUMA nodes, allowed CPUs/memory nodes of this container, GPU <-> NUMA affinity
nvidia-smi topo -m 2>&1 | head -20
grep -E "Cpus_allowed_list|Mems_allowed_list" /proc/self/status
ls /sys/devices/system/node/ | grep node
for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) $(grep MemTotal $n/meminfo); done
for g in /sys/bus/pci/drivers/nvidia/0000*; do echo $g $(cat $g/numa_node) $(cat $g/local_cpulist); done 2>/dev/null
nproc; free -g | head -2
