#!/usr/bin/env python
"""L2 -> SM operand traffic of every launch in an .ncu-rep: TMA load bytes, rate, LTS throughput, tensor-pipe activity."""
import csv, subprocess, sys
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        g = lambda n: (v[h.index(n)], u[h.index(n)]) if n in h else ("-", "")
        print(rep.split("/")[-1])
        for n in ["gpu__time_duration.sum", "sm__cycles_elapsed.max.per_second", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
                  "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum.per_second", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
                  "l1tex__m_l1tex2xbar_write_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                  "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum",
                  "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum"]:
            print("   %-75s %s %s" % (n, *g(n)))
