import json,sys
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): 
        if line: print(line[:200])
        continue
    d=json.loads(line)
    print(f"{d['config'].get('variant')} rays/s={d['value']/1e6:.2f}M step={d['ms_per_step']:.2f}ms fwd={d['roofline_fwd']['kernel_ms']:.2f} bwd={d.get('roofline_bwd', d['roofline'])['kernel_ms']:.2f} frac_bwd={d.get('roofline_bwd', d['roofline'])['frac']:.4f} dominant={d['roofline']['kernel']} frac_fwd={d['roofline_fwd']['frac']:.4f} e2e={d['e2e']['value']/1e6:.2f}M clk={d['clocks']['sm_mhz']}")
