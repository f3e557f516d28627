"""CPU: the map-producing CNNs (map_producers.py) against the reference's own modules when the reference tree is present
(this container; it does not travel to the GPU box): same parameter names (its checkpoints load with load_state_dict) and
the same outputs for the same weights.  Always: output layout and normalisation of MapProducer.produce."""
import os
import sys

import pytest
import torch

from mpp_cnn_rs_object_detection_b200.map_producers import MapProducer

REF = "/root/reference"


def test_produce_layout_and_normalisation():
    torch.manual_seed(0)
    mp = MapProducer()
    det, marks = mp.produce(torch.rand(2, 3, 70, 90))  # 70 x 90: not a multiple of 8 -> padded and cropped (unet.py:9-21)
    assert det.shape == (2, 70, 90) and marks.shape == (2, 3, 70, 90, 32)
    assert det.dtype == torch.float32 and marks.dtype == torch.float32 and marks.is_contiguous()
    assert float(det.min()) > 0 and float(det.max()) < 1
    torch.testing.assert_close(marks.sum(-1), torch.ones(2, 3, 70, 90), rtol=0, atol=1e-5)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models", "position_net")), reason="reference tree not present")
def test_same_parameters_and_outputs_as_the_reference_modules():
    sys.path.insert(0, REF)
    try:
        from models.position_net.pos_net import PosNet
        from models.position_net.torch_div import Divergence
        from models.shape_net.shape_net import ShapeNet as RefShapeNet
    finally:
        sys.path.remove(REF)
    torch.manual_seed(1)
    cpu = torch.device("cpu")
    ref_pos = PosNet(3, 3, cpu, hidden_dims=[32, 64, 128, 256]).eval()
    ref_shape = RefShapeNet(3, 3, 32, cpu, hidden_dims=[32, 64, 128, 256]).eval()
    mp = MapProducer().eval()
    assert set(ref_pos.state_dict()) == set(mp.posnet.state_dict())
    assert set(ref_shape.state_dict()) == set(mp.shapenet.state_dict())
    mp.posnet.load_state_dict(ref_pos.state_dict())
    mp.shapenet.load_state_dict(ref_shape.state_dict())
    x = torch.rand(1, 3, 64, 96)
    with torch.no_grad():
        out = ref_pos(x)
        torch.testing.assert_close(mp.posnet(x), out, rtol=1e-5, atol=1e-5)
        for a, b in zip(mp.shapenet(x), ref_shape(x)):
            torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)
        # detection map: sigmoid(div_clf(Divergence(vec) * mask))  (pos_net_model.py:75-79,338-349)
        div = Divergence(div_channels=[0, 1], mask_channel=2)
        want_det = torch.sigmoid(mp.div_clf(div(out))).squeeze()
        det, marks = mp.produce(x)
        torch.testing.assert_close(det[0], want_det, rtol=1e-5, atol=1e-5)
        want_marks = [torch.softmax(t, dim=1)[0].permute(1, 2, 0) for t in ref_shape(x)]
        for i in range(3):
            torch.testing.assert_close(marks[0, i], want_marks[i], rtol=1e-5, atol=1e-5)
