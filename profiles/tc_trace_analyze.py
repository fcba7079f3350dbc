import json, collections
d=json.load(open('gpurun_out/tc_trace.json'))
comp=[(t,tag) for tag,t in d if tag<100]; mma=[(t,tag-100) for tag,t in d if tag>=100]
names={1:"stage start",2:"D1 ready",3:"tmem ld done",4:"split done",5:"tmem st done",6:"arrived",20:"tile start",21:"q1 done+fusion",22:"g net done"}
def show(ev,names):
    agg=collections.defaultdict(list)
    for (ta,a),(tb,b) in zip(ev[:-1],ev[1:]):
        agg[(a,b)].append(tb-ta)
    for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
        v2=sorted(v)
        print("  %-16s -> %-16s n=%4d mean=%7.0f med=%6d min=%6d max=%6d total=%8d"%(names.get(k[0],k[0]),names.get(k[1],k[1]),len(v),sum(v)/len(v),v2[len(v)//2],v2[0],v2[-1],sum(v)))
show(comp,names)
n2={10:"wait A1",11:"A1 ok",12:"stage staged",13:"L2/L1 issued"}
print("MMA thread"); show(mma,n2)
tiles=[t for t,tag in comp if tag==20]
print("tiles", len(tiles), [tiles[i+1]-tiles[i] for i in range(len(tiles)-1)], "span", comp[-1][0]-comp[0][0])
t0=tiles[1]; t1=tiles[2]
ev=sorted([(t-t0,"C",names[tag]) for t,tag in comp if t0<=t<t1]+[(t-t0,"M",n2[tag]) for t,tag in mma if t0<=t<t1])
for e in ev[:60]: print("%7d %s %s"%e)
