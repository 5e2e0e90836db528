#!/bin/bash
# usage: tools/ncu_summarize.sh <report.ncu-rep> <mangled-kernel-substring> <out.txt>   (run in the build container: no GPU needed)
set -e
REP=$1; KERN=$2; OUT=$3
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
( cd $TMP && cuobjdump -xelf all $ROOT/minimal_volumetric_path_tracer_b200/libvpt_b200.so >/dev/null 2>&1 && nvdisasm -g -c vpt_kernels_f32.sm_100a.cubin > f32.sass 2>/dev/null )
ncu -i $REP --page source --csv > $TMP/src.csv 2>/dev/null
{
  echo "# ncu summary of $(basename $REP) (kernel *$KERN*), produced by tools/ncu_summarize.sh"
  echo "# capture: ncu --set full --clock-control none --import-source on (B200, sm_100a); per-launch values"
  echo
  ncu -i $REP --page details 2>/dev/null | grep -E "^\s+(Duration|Elapsed Cycles|SM Frequency|Compute \(SM\) Throughput|Memory Throughput|DRAM Throughput|Executed Ipc Active|Issue Slots Busy|SM Busy|No Eligible|One or More Eligible|Active Warps Per Scheduler|Eligible Warps Per Scheduler|Warp Cycles Per Issued|Avg. Active Threads Per Warp|Avg. Not Predicated|Registers Per Thread|Theoretical Occupancy|Achieved Occupancy|Branch Efficiency|L1/TEX Hit Rate|L2 Hit Rate|Block Limit Registers|Block Limit Shared)"
  echo
  echo "## raw metrics"
  ncu -i $REP --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]; r=rows[-1]
want=['dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','sm__inst_executed.sum','smsp__thread_inst_executed.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_xu.sum','sm__inst_executed_pipe_fp64.sum','smsp__inst_executed_pipe_fma.sum','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.sum','sm__inst_executed_pipe_fmalite.sum','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active']
for i,n in enumerate(h):
    if n in want: print('%-75s %s %s'%(n, r[i], rows[1][i] if len(rows)>2 else ''))
"
  echo
  echo "## stall reasons (share of warp-state samples) and opcode mix (share of executed warp-instructions)"
  python3 - $TMP/src.csv <<'PY'
import csv,sys,re,collections
rows=list(csv.reader(open(sys.argv[1])))
hdr=next(i for i,r in enumerate(rows) if r and r[0]=="Address")
H=rows[hdr]; data=rows[hdr+1:]
tot={n:sum(int(r[j] or 0) for r in data) for j,n in enumerate(H) if n.startswith('stall_') and 'Not Issued' not in n}
s=sum(tot.values())
print("  ".join("%s %.1f%%"%(k[6:],100*v/s) for k,v in sorted(tot.items(), key=lambda x:-x[1])[:9]))
ops=collections.Counter(); ci=H.index("Instructions Executed")
for r in data:
    m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[1]); ops[m.group(2).split('.')[0] if m else '?']+=int(r[ci])
t=sum(ops.values()); print("  ".join("%s %.1f%%"%(k,100*v/t) for k,v in ops.most_common(16)))
print("SASS instructions in kernel: %d (%.1f KB)"%(len(data), len(data)*16/1024))
PY
  echo
  echo "## executed instructions by source line (inst% = share of warp-instructions, lanes = avg active threads)"
  python3 $ROOT/tools/ncu_by_line.py $TMP/src.csv $TMP/f32.sass $KERN 40
} > $OUT
rm -rf $TMP
echo wrote $OUT
