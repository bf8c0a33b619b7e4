cd scripts/probes
./scan_probe 2>&1 | tail -14
M=dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum
for i in 10 18 21 24 25 26 27 28 29 30 31 32; do ncu --metrics $M -c 1 ./scan_probe $i 2>&1 | grep -E "^ *[0-9]+ [a-z]|dram__bytes_read|lts__t_sectors|gpu__time" | tr '\n' ' '; echo; done
