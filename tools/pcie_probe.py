"""tools/pcie_probe.py — pinned host <-> device copy rates of this box for the byte counts of bench.py's e2e leg (one copy against the
number of separate arrays dml_upload / dml_download move): the floor the host-buffer path is compared with in DESIGN.md §6."""
import torch, time
n=109647
for total,parts,name in ((18420696,1,"D2H one"),(18420696,11,"D2H 11 parts"),(14911992,1,"H2D one"),(14911992,9,"H2D 9 parts")):
    d=torch.empty(total,dtype=torch.uint8,device="cuda"); h=torch.empty(total,dtype=torch.uint8,pin_memory=True)
    ch=total//parts
    for rep in range(3):
        torch.cuda.synchronize(); t0=time.perf_counter()
        for k in range(20):
            for p in range(parts):
                a,b=p*ch,(p+1)*ch if p<parts-1 else total
                if name.startswith("D2H"): h[a:b].copy_(d[a:b],non_blocking=True)
                else: d[a:b].copy_(h[a:b],non_blocking=True)
            torch.cuda.synchronize()
        dt=(time.perf_counter()-t0)/20
    print(name, "%.3f ms  %.1f GB/s"%(dt*1e3,total/dt/1e9))
