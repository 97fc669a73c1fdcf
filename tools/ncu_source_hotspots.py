"""Per-source-line instruction and stall-sample breakdown of one kernel from an ncu report with imported source
(usage: python tools/ncu_source_hotspots.py <kernel name> <sites in the capture> <output file>; reads gpurun_out/r1_full_v16.ncu-rep)."""
import csv, sys, subprocess
kname, sites, outp = sys.argv[1], float(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu","-i","gpurun_out/r1_full_v16.ncu-rep","--page","source","--print-source","cuda,sass","--csv","--kernel-name",kname],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=None; cur_file=None; agg={}
for r in rows:
    if len(r)==2 and r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if len(r)>10 and r[0]=='Line No': hdr=r; continue
    if hdr and len(r)==len(hdr) and r[0]!='':
        d=dict(zip(hdr,r))
        try: ie=int(d['Instructions Executed']); sm=int(d['# Samples'])
        except: continue
        a=agg.setdefault((cur_file,int(r[0])),[0,0,r[1]]); a[0]+=ie; a[1]+=sm
tot_i=sum(v[0] for v in agg.values()); tot_s=sum(v[1] for v in agg.values())
with open(outp,"w") as f:
    f.write("# %s: executed warp instructions and warp-stall samples per source line\n" % kname)
    f.write("# source: gpurun_out/r1_full_v16.ncu-rep (ncu --set full --clock-control none --import-source on, tools/gpu_probe.py 1e7 0:\n")
    f.write("#         10 Mb contig, 97,822 sites), read with: ncu -i r1_full_v16.ncu-rep --page source --print-source cuda,sass --csv --kernel-name %s\n" % kname)
    f.write("# totals: %d warp instructions (%.0f per site), %d stall samples.\n\n" % (tot_i, tot_i/sites, tot_s))
    f.write("## top 30 lines by executed instructions\n")
    for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:30]:
        f.write("%-22s %5d  %5.1f %% inst  %5.1f %% samples  %s\n"%(k[0],k[1],100*v[0]/tot_i,100*v[1]/tot_s,v[2][:110]))
    f.write("\n## top 25 lines by stall samples\n")
    for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]:
        f.write("%-22s %5d  %5.1f %% samples  %5.1f %% inst  %s\n"%(k[0],k[1],100*v[1]/tot_s,100*v[0]/tot_i,v[2][:110]))
