import sys, torch
sys.path.insert(0, '.')
import bench, detectron2_b200.layers as L
d = bench.make_image_inputs(0)
dev = torch.device('cuda')
masks, boxes = d["masks"].to(dev), d["proposals"][:100].to(dev)
def t(fn, rep=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(rep): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / rep * 1e3
byte_us = t(lambda: L.paste_masks_in_image(masks, boxes, (800, 1333), 0.5))
pack_us = t(lambda: L.paste_masks_in_image_packed(masks, boxes, (800, 1333), 0.5))
hb = torch.empty((100, 800, 1333), dtype=torch.bool).pin_memory(); hp = torch.empty((100, 800, 42), dtype=torch.int32).pin_memory()
e2e_byte = t(lambda: hb.copy_(L.paste_masks_in_image(masks, boxes, (800, 1333), 0.5), non_blocking=True), 20)
e2e_pack = t(lambda: hp.copy_(L.paste_masks_in_image_packed(masks, boxes, (800, 1333), 0.5), non_blocking=True), 20)
print({"paste_byte_us": round(byte_us, 1), "paste_packed_us": round(pack_us, 1), "paste_plus_d2h_byte_us": round(e2e_byte, 1), "paste_plus_d2h_packed_us": round(e2e_pack, 1)})
