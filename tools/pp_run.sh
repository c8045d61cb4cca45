echo base; python tools/model_probe.py
for v in 4 5; do echo variant $v; POSG_B200_LIB=tools/variants/libposg_g$v.so python tools/model_probe.py; done
