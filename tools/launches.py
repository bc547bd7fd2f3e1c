"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of
one decode step (between two decode_prepare launches) and of one prefill step."""
import collections, csv, re, sys
def load(path):
    lines=[l for l in open(path) if not l.startswith('==')]
    rows=[]
    for row in csv.DictReader(lines):
        try: rows.append((row['Kernel Name'], float(row['Metric Value'].replace(',','')), row.get('Grid Size','')))
        except Exception: pass
    return rows
def summarize(step, title):
    tot=sum(d for _,d,_ in step)
    print("%s: %d launches, %.1f us total"%(title,len(step),tot/1e3))
    agg={}
    for name,d,g in step:
        key=re.sub(r'\(.*','',name).replace('void llmlb::','').replace('llmlb::','')+' grid='+g
        a=agg.setdefault(key,[0,0.0]); a[0]+=1; a[1]+=d
    for k,(n,d) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
        print("  %4d x avg %8.2f us = %8.1f us  %5.1f%%  %s"%(n,d/n/1e3,d/1e3,100*d/tot,k[:100]))
if __name__=="__main__":
    rows=load(sys.argv[1])
    idx=[i for i,x in enumerate(rows) if 'decode_prepare' in x[0]]
    if len(idx)>=2: summarize(rows[idx[-2]:idx[-1]], "decode step")
    pi=[i for i,x in enumerate(rows) if 'slot_init' in x[0]]
    if pi:
        e=[i for i in idx if i>pi[-1]]
        summarize(rows[pi[-1]:(e[0] if e else len(rows))], "prefill step")
