"""Workload tables vs the dimensions SURVEY.md section 8 records from the
reference's own model files (and scripts/dp_block_partition.py:92-93)."""
from kfac_pytorch_b200.workloads import (algorithmic_work, bert_base_qa_layers, layer_dims,
                                         resnet50_layers, rows_of, vgg16_cifar_layers)

R50_DA = [147,64,576,64,64,256,576,64,256,576,64,256,1152,128,256,512,1152,128,512,1152,128,512,1152,128,512,2304,256,512,1024,2304,256,1024,2304,256,1024,2304,256,1024,2304,256,1024,2304,256,1024,4608,512,1024,2048,4608,512,2048,4608,512,2049]
R50_DG = [64,64,64,256,256,64,64,256,64,64,256,128,128,512,512,128,128,512,128,128,512,128,128,512,256,256,1024,1024,256,256,1024,256,256,1024,256,256,1024,256,256,1024,256,256,1024,512,512,2048,2048,512,512,2048,512,512,2048,1000]


def test_resnet50_table():
    ls = resnet50_layers()
    assert [layer_dims(l)[0] for l in ls] == R50_DA
    assert [layer_dims(l)[1] for l in ls] == R50_DG
    assert rows_of(ls[0], 32) == 401408 and rows_of(ls[1], 32) == 100352 and rows_of(ls[-2], 32) == 1568
    w = algorithmic_work(ls, 32, 'eigen_dp')
    assert abs(w['F_factor'] / 1e9 - 1605) < 5
    assert abs(w['F_inverse'] / 1e9 - 4075) < 10
    assert abs(w['F_apply'] / 1e9 - 312) < 2


def test_vgg16_and_bert_tables():
    v = vgg16_cifar_layers()
    assert [layer_dims(l)[0] for l in v] == [28,577,577,1153,1153,2305,2305,2305,4609,4609,4609,4609,4609,513]
    assert [layer_dims(l)[1] for l in v] == [64,64,128,128,256,256,256,512,512,512,512,512,512,10]
    assert [rows_of(l, 128) for l in v][:3] == [131072, 131072, 32768]
    b = bert_base_qa_layers()
    assert len(b) == 73
    dims = [layer_dims(l) for l in b]
    assert dims.count((769, 768)) == 48 and dims.count((769, 3072)) == 12 and dims.count((3073, 768)) == 12
    assert dims[-1] == (769, 2)
