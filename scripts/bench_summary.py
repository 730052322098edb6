import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"])
sm=d["extra"].get("split_merge",{})
print(sm.get("error"), sm.get("total_ms"), sm.get("route"), sm.get("nccl_two_sided_total_ms"))
for l in sm.get("legs",[]): print(" leg", l.get("transform"), l.get("total_ms"), l.get("bitwise_equal"), l.get("error"))
for k,v in sm.get("stft_variants",{}).items():
    print(" var", k, v.get("total_ms"), v.get("bitwise_equal")) if isinstance(v,dict) else print(k,v)
print([ (c["transform"], round(c.get("ms_per_step",0),2), c.get("error")) for c in d["extra"]["configs"]])
print([ (c["transform"], round(c["ms_per_step"],2)) for c in d["extra"]["dct"]])
c=d["extra"]["device_chain"]; print(c["ms_per_batch"], c["onesided"]["ms_per_batch"]) if c else None
co=d["extra"].get("c_order"); print("c_order", [ (c["transform"], round(c.get("ms_per_step",0),2), round(c.get("roofline",{}).get("frac",0),3), c.get("parity_max_rel_err"), c.get("error")) for c in co] if isinstance(co,list) else co)
