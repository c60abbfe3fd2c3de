for c in 0 1; do for d in 0 1; do
CG_ENABLE_2CTA=$c CG_TC_DBG=$d ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none -k regex:conv_tc --csv --log-file gpurun_out/x2_c${c}_d${d}.csv python tools/prof_conv.py 32 3 > gpurun_out/x2_c${c}_d${d}.log 2>&1
done; done
python - <<'P'
import csv,glob
for f in sorted(glob.glob('gpurun_out/x2_c*_d*.csv')):
    rows=[r for r in csv.reader(open(f)) if len(r)>10]
    h=rows[0]; out=[]
    for r in rows[1:]:
        d=dict(zip(h,r))
        if d['Metric Name']=='gpu__time_duration.sum': out.append((d['Kernel Name'][:24], d['Metric Value']))
    print(f, out)
P
