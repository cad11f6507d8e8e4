#!/bin/bash
# what the GPU box has (GL/EGL libraries for SURVEY 8f#4, CPU / NUMA topology for the host-fed path)
mkdir -p gpurun_out
{
echo "== nproc / lscpu"; nproc; lscpu | egrep 'Model name|Socket|Core|Thread|NUMA|L3|CPU\(s\)'
echo "== affinity"; taskset -p $$; cat /sys/fs/cgroup/cpu.max 2>/dev/null
echo "== numa"; ls /sys/devices/system/node/ 2>/dev/null; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) $(grep MemTotal $n/meminfo); done
echo "== nvidia-smi topo"; nvidia-smi topo -m 2>&1 | head -30
echo "== gpu numa"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo $d $(cat $d/class) numa=$(cat $d/numa_node) local_cpus=$(cat $d/local_cpulist); fi; done
echo "== GL / EGL / OpenCL libraries"; ldconfig -p | egrep -i 'egl|libGL|opengl|opencl|glx|gbm' ; ls /usr/lib/x86_64-linux-gnu | egrep -i 'nvidia|egl|libGL' | head -40
echo "== NVIDIA_DRIVER_CAPABILITIES=$NVIDIA_DRIVER_CAPABILITIES"; ls /usr/share/glvnd/egl_vendor.d /etc/OpenCL/vendors 2>&1
echo "== meminfo"; head -3 /proc/meminfo; ulimit -l
echo "== host feed probe"; tools/bin/host_feed_probe
} > gpurun_out/box_probe.txt 2>&1
tail -40 gpurun_out/box_probe.txt
