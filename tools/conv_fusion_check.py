import sys, torch
sys.path.insert(0,'detr-object-detection_b200')
from detr_b200.harness import _Backbone
torch.manual_seed(0)
b=_Backbone('resnet50').cuda().to(memory_format=torch.channels_last)
for n,buf in b.named_buffers():
    if 'running_var' in n: buf.uniform_(0.5,1.5)
    elif 'running_mean' in n: buf.normal_(0,0.1)
    elif n.endswith('weight'): buf.uniform_(0.5,1.5)
    elif n.endswith('bias'): buf.normal_(0,0.1)
x=torch.randn(2,3,256,320,device='cuda').contiguous(memory_format=torch.channels_last)
res={}
for fused in (True, False):
    b.fuse_relu=fused
    b.zero_grad()
    with torch.autocast('cuda',dtype=torch.bfloat16):
        y=b(x)
    y.float().square().mean().backward()
    res[fused]=(y.float().clone(), b.backbone.layer2[0].conv1.weight.grad.clone(), b.backbone.conv1.weight.grad.clone())
for i,name in enumerate(['out','grad layer2.0.conv1','grad conv1']):
    a,c=res[True][i],res[False][i]
    print(name, (a-c).abs().max().item()/ (c.abs().max().item()+1e-12))
