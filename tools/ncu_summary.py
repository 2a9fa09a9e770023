#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + top stall locations.  usage: ncu_summary.py rep [topN]"""
import csv, subprocess, sys, collections, io
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 25
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw))); hdr,units=rows[0],rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__cycles_elapsed.avg','sm__cycles_elapsed.avg.per_second','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__waves_per_multiprocessor','lts__t_bytes.sum','l1tex__t_bytes.sum','lts__t_sector_hit_rate.pct','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__thread_inst_executed_per_inst_executed.ratio']
for r in rows[2:]:
    print('====', r[hdr.index('Kernel Name')][:80])
    for i,h in enumerate(hdr):
        if h in want: print(f"  {h:75s} {units[i]:10s} {r[i]}")
    st={h:float(r[i]) for i,h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and r[i]}
    print("  stalls/issue:", ", ".join(f"{k.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for k,v in sorted(st.items(),key=lambda x:-x[1])[:7]))
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
hdr=None; out=[]; k=0
for r in rows:
    if r and r[0]=="Kernel Name": k+=1; continue
    if r and r[0]=="Address": hdr=r; continue
    if hdr is None or k!=1: continue
    out.append(r)
if hdr:
    ia=hdr.index("Source"); isamp=hdr.index("# Samples"); iex=hdr.index("Instructions Executed")
    stalls=[i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot=sum(int(r[isamp] or 0) for r in out) or 1
    print("total samples",tot,"sass instrs",len(out))
    for idx,r in sorted(enumerate(out),key=lambda x:-int(x[1][isamp] or 0))[:topn]:
        s=int(r[isamp] or 0)
        st=sorted(((hdr[i],int(r[i] or 0)) for i in stalls if (r[i] or '0')!='0'),key=lambda x:-x[1])[:2]
        print(f"{s:6d} {100*s/tot:5.1f}% #{idx:5d} ex={r[iex]:>8s} {r[ia][:64]:64s} {st}")
