for i in 1 2; do for v in old new new2; do
  if [ $v = new ]; then L=qibotn_b200/libqibotn_b200.so; else L=_ab/lib$v.so; fi
  QTN_LIB=$PWD/$L timeout 150 python scripts/tc_pair_probe.py time > gpurun_out/ab_${v}_$i.jsonl 2>> gpurun_out/ab.err
done; done
python - <<PY
import json,glob
res={}
for f in sorted(glob.glob("gpurun_out/ab_*_*.jsonl")):
    v=f.split("ab_")[1].rsplit("_",1)[0]
    for l in open(f):
        d=json.loads(l); key=(d["m"],d["n"],d["k"])
        res.setdefault(key,{}).setdefault(v,[]).append((d["single"]["ms"],d["pair"]["ms"]))
for key,vs in res.items():
    print(key, " | ".join("%s single %s pair %s"%(v," ".join("%.3f"%x[0] for x in r)," ".join("%.3f"%x[1] for x in r)) for v,r in vs.items()))
PY
tail -c 300 gpurun_out/ab.err
