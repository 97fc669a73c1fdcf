import sys, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from nanocaller_b200.host import weights as W
from oracle import cnn_oracle, snp_oracle
from tests.golden_util import golden_chunk, load_case

def q_bf16(t): return t.to(torch.bfloat16).to(torch.float32)
def q_tf32(t):
    i = t.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF   # round to nearest (ties away) 10-bit mantissa
    return i.view(torch.float32)
def q_split(t):  # bf16 hi + bf16 lo  (effectively ~16 bit mantissa)
    hi = q_bf16(t); lo = q_bf16(t-hi); return hi+lo

def run(w, x, ref, qa, qw, layers=('conv1','conv2','conv3','fc1')):
    dt=torch.float32
    def conv(x, name, stride, same, q):
        k=torch.as_tensor(w[name+'/kernel']).permute(3,2,0,1).contiguous(); b=torch.as_tensor(w[name+'/bias'])
        pad=(k.shape[2]//2,k.shape[3]//2) if same else 0
        if q: x=qa(x); k=qw(k)
        return F.selu(F.conv2d(x.double(),k.double(),b.double(),stride=stride,padding=pad).float())
    x=torch.as_tensor(x).permute(0,3,1,2).contiguous()
    c1=torch.cat([conv(x,'conv1_1',1,True,'conv1' in layers),conv(x,'conv1_2',1,True,'conv1' in layers),conv(x,'conv1_3',1,True,'conv1' in layers)],1)
    c2=conv(c1,'conv2',(1,2),False,'conv2' in layers)
    c3=conv(c2,'conv3',(1,2),False,'conv3' in layers)
    flat=c3.permute(0,2,3,1).reshape(c3.shape[0],-1)
    k=torch.as_tensor(w['fc1/kernel']); b=torch.as_tensor(w['fc1/bias'])
    if 'fc1' in layers: flat=qa(flat); k=qw(k)
    fc1=F.selu((flat.double()@k.double()+b.double()).float())
    fa=F.selu(fc1@torch.as_tensor(w['fa/kernel'])+torch.as_tensor(w['fa/bias']))
    ref=torch.as_tensor(ref)
    outs=[]
    for j,bb in enumerate('AGTC'):
        z=torch.cat([fa,ref[:,j:j+1]],1)@torch.as_tensor(w[bb+'/kernel'])+torch.as_tensor(w[bb+'/bias'])
        outs.append(torch.softmax(z,-1)[:,1])
    return torch.stack(outs,1).numpy()

ident=lambda t:t
for model in ['ONT-HG002','CCS-HG002','NanoCaller1','ONT-HG001']:
    tensors, meta = W.load_model('snp', model)
    rs,dct,chunks,bed,g=load_case('ont_diploid')
    xs=[];refs=[]
    for ci in range(len(chunks)):
        w_=golden_chunk(g,ci); xs.append(snp_oracle.scale_counts(w_['mat'], meta['train_coverage'] or 30., coverage=float(w_['depth']))); refs.append(w_['ref'].astype(np.float32))
    x=np.concatenate(xs); ref=np.concatenate(refs)
    base=run(tensors,x,ref,ident,ident,())
    o=cnn_oracle.snp_probs(tensors,x,ref)
    print(model,'n',len(x),'fp32-oracle vs fp64-acc', np.abs(o-base).max())
    for name,qa,qw in [('bf16/bf16',q_bf16,q_bf16),('tf32/tf32',q_tf32,q_tf32),('act bf16, w split',q_bf16,q_split),('act split, w bf16',q_split,q_bf16),('split/split',q_split,q_split),('act tf32,w fp32',q_tf32,ident)]:
        e=np.abs(run(tensors,x,ref,qa,qw)-base); print('   %-20s max %.2e  p99.9 %.2e mean %.2e'%(name,e.max(),np.quantile(e,0.999),e.mean()))
    for L in ['conv1','conv2','conv3','fc1']:
        e=np.abs(run(tensors,x,ref,q_bf16,q_bf16,(L,))-base); print('   bf16 only in %-6s max %.2e'%(L,e.max()))

print("==== 3-term emulation")
def split16(t, dt):
    hi = t.to(dt).to(torch.float32); lo = (t-hi).to(dt).to(torch.float32); return hi, lo
def run3(w, x, ref, dt, terms=3):
    stats={}
    def mm3(conv_fn, a, k):
        ah,al=split16(a,dt); kh,kl=split16(k,dt)
        r=conv_fn(ah.double(),kh.double())+conv_fn(ah.double(),kl.double())+conv_fn(al.double(),kh.double())
        if terms==4: r=r+conv_fn(al.double(),kl.double())
        return r
    def conv(x, name, stride, same):
        k=torch.as_tensor(w[name+'/kernel']).permute(3,2,0,1).contiguous(); b=torch.as_tensor(w[name+'/bias'])
        pad=(k.shape[2]//2,k.shape[3]//2) if same else 0
        r=mm3(lambda a,kk:F.conv2d(a,kk,None,stride=stride,padding=pad), x, k)+b.double().view(1,-1,1,1)
        return F.selu(r.float())
    x=torch.as_tensor(x).permute(0,3,1,2).contiguous()
    c1=torch.cat([conv(x,'conv1_1',1,True),conv(x,'conv1_2',1,True),conv(x,'conv1_3',1,True)],1)
    c2=conv(c1,'conv2',(1,2),False); c3=conv(c2,'conv3',(1,2),False)
    stats=dict(x=x.abs().max().item(),c1=c1.abs().max().item(),c2=c2.abs().max().item(),c3=c3.abs().max().item())
    flat=c3.permute(0,2,3,1).reshape(c3.shape[0],-1)
    k=torch.as_tensor(w['fc1/kernel']); b=torch.as_tensor(w['fc1/bias'])
    fc1=F.selu((mm3(lambda a,kk:a@kk, flat,k)+b.double()).float())
    stats['fc1']=fc1.abs().max().item()
    fa=F.selu(fc1@torch.as_tensor(w['fa/kernel'])+torch.as_tensor(w['fa/bias']))
    ref=torch.as_tensor(ref); outs=[]
    for j,bb in enumerate('AGTC'):
        z=torch.cat([fa,ref[:,j:j+1]],1)@torch.as_tensor(w[bb+'/kernel'])+torch.as_tensor(w[bb+'/bias'])
        outs.append(torch.softmax(z,-1)[:,1])
    return torch.stack(outs,1).numpy(), stats
for model in ['ONT-HG002','CCS-HG002','NanoCaller1','ONT-HG001','CLR-HG002']:
    tensors, meta = W.load_model('snp', model)
    rs,dct,chunks,bed,g=load_case('ont_diploid')
    xs=[];refs=[]
    for ci in range(len(chunks)):
        w_=golden_chunk(g,ci); xs.append(snp_oracle.scale_counts(w_['mat'], meta['train_coverage'] or 30., coverage=float(w_['depth']))); refs.append(w_['ref'].astype(np.float32))
    x=np.concatenate(xs); ref=np.concatenate(refs)
    base=run(tensors,x,ref,ident,ident,())
    for nm,dt,terms in [('bf16 3-term',torch.bfloat16,3),('bf16 4-term',torch.bfloat16,4),('fp16 3-term',torch.float16,3)]:
        o,st=run3(tensors,x,ref,dt,terms); e=np.abs(o-base)
        print(model,'%-12s max %.2e mean %.2e'%(nm,e.max(),e.mean()), {k:round(v,1) for k,v in st.items()} if dt==torch.float16 else '')
    wmax=max(np.abs(v).max() for v in tensors.values()); print('   max |w|',wmax)
