"""Summary of one ncu report: headline metrics + the source lines with the most stall samples.
``python tools/ncu_lines.py report.ncu-rep [top]`` (needs ncu on PATH; reports are made with --set full --import-source on)."""
import csv,subprocess,io,sys
rep=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 30
out=subprocess.run(f"ncu -i {rep} --page raw --csv",shell=True,capture_output=True,text=True).stdout
r=list(csv.reader(io.StringIO(out))); h,u,v=r[0],r[1],r[2]
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed_op_shared_atom.sum','lts__t_bytes.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum']
for i,k in enumerate(h):
    if k in keys or ('issue_stalled' in k and 'per_issue_active' in k and float(v[i] or 0)>0.3): print(k,u[i],v[i])
out=subprocess.run(f"ncu -i {rep} --page source --print-source sass,cuda --csv",shell=True,capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
data=[]
for r in rows:
    if len(r)>8 and r[0] not in ("","Line No") and r[2]=='-':
        try: data.append((int(r[6] or 0), int(r[7] or 0), r[0], r[1][:100]))
        except Exception as e: pass
tot_s=sum(d[0] for d in data); tot_i=sum(d[1] for d in data)
print(tot_s,tot_i)
for d in sorted(data,key=lambda x:-x[0])[:top]: print(f"{d[0]/tot_s*100:5.1f}% smp {d[1]/tot_i*100:5.1f}% inst  {d[2]} {d[3]}")
