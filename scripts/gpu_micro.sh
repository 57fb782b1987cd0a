cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
{
./scripts/micro/gather_bench all
M="dram__bytes_read.sum,dram__sectors_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"
ncu --metrics $M --clock-control none -k regex:walk_kernel ./scripts/micro/gather_bench walk 2>&1 | grep -E "walk_kernel|dram__|lts__|gpu__time|l1tex__" 
} 2>&1 | tee gpurun_out/micro1.txt
