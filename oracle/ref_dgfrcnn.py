"""CPU restatement of the reference's DGFRCNN training step on STOCK torchvision ops.
TEST INFRASTRUCTURE / bench reference arm only (see oracle/__init__.py).

What it restates (the reference itself is Python and cannot travel to the GPU box):
  * fasterrcnn.py:90-196   RegionProposalNetworkWILDS  — torchvision RPN with per-image losses
  * fasterrcnn.py:198-305  fastrcnn_loss / RoIHeadsWILDS — per-image losses, labels passed to the
                           box head, post-processing also in training
  * fasterrcnn.py:331-499  TwoMLPHead(x, box_labels), FastWILDS (zero mean / unit std transform)
  * DGcommon.py:33-113     GRLayer, ImageDAFPN
  * DGFRCNN.py:4-201       instance heads, hooks, 5-mode training_step, SGD
Everything below the per-image loss plumbing is the installed torchvision CPU code path
(RegionProposalNetwork.filter_proposals / assign_targets_to_anchors, RoIHeads.select_training_samples /
postprocess_detections, MultiScaleRoIAlign, nms, roi_align) — i.e. the reference's arithmetic.
It is pinned against the real reference modules by oracle/gen_golden.py + tests/test_oracle_golden.py.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F
from torch import Tensor, nn
from torchvision.models.detection.backbone_utils import resnet_fpn_backbone
from torchvision.models.detection.faster_rcnn import FastRCNNPredictor
from torchvision.models.detection.anchor_utils import AnchorGenerator
from torchvision.models.detection.generalized_rcnn import GeneralizedRCNN
from torchvision.models.detection.roi_heads import RoIHeads
from torchvision.models.detection.rpn import RegionProposalNetwork, RPNHead
from torchvision.models.detection.transform import GeneralizedRCNNTransform
from torchvision.ops import MultiScaleRoIAlign


class PerImageRPN(RegionProposalNetwork):
    """Stock RPN; only the loss reduction differs: one value per image (fasterrcnn.py:105-140)."""

    def compute_loss(self, objectness, pred_bbox_deltas, labels, regression_targets):
        n_img = len(labels)
        obj = objectness.reshape(n_img, -1)              # image-major after concat_box_prediction_layers
        dl = pred_bbox_deltas.reshape(n_img, -1, 4)
        l_obj, l_box = [], []
        for i in range(n_img):
            pos_m, neg_m = self.fg_bg_sampler([labels[i]])               # two randperms per image
            pos = torch.where(pos_m[0])[0]
            neg = torch.where(neg_m[0])[0]
            both = torch.cat([pos, neg])
            l_box.append(F.smooth_l1_loss(dl[i][pos], regression_targets[i][pos], beta=1 / 9, reduction="sum")
                         / both.numel())
            l_obj.append(F.binary_cross_entropy_with_logits(obj[i][both], labels[i][both]))
        return torch.stack(l_obj), torch.stack(l_box)


class PerImageRoIHeads(RoIHeads):
    """Stock RoI heads with fasterrcnn.py:247-305's differences."""

    def forward(self, features, proposals, image_shapes, targets=None):
        labels = regression_targets = None
        if self.training:
            proposals, _, labels, regression_targets = self.select_training_samples(proposals, targets)
        pooled = self.box_roi_pool(features, proposals, image_shapes)
        box_features = self.box_head(pooled, labels)                     # labels reach the hook (DGFRCNN.py:89-91)
        class_logits, box_regression = self.box_predictor(box_features)
        losses = {}
        if self.training:
            per = [len(l) for l in labels]                                # 512 each (fasterrcnn.py:211-212)
            l_cls, l_box = [], []
            for lg, rg, lab, tgt in zip(class_logits.split(per), box_regression.split(per), labels, regression_targets):
                l_cls.append(F.cross_entropy(lg, lab))
                pos = torch.where(lab > 0)[0]
                rg = rg.reshape(lg.shape[0], -1, 4)
                l_box.append(F.smooth_l1_loss(rg[pos, lab[pos]], tgt[pos], beta=1 / 9, reduction="sum") / lab.numel())
            losses = {"loss_classifier": torch.stack(l_cls), "loss_box_reg": torch.stack(l_box)}
        boxes, scores, labs = self.postprocess_detections(class_logits, box_regression, proposals, image_shapes)
        return [{"boxes": b, "labels": l, "scores": s} for b, l, s in zip(boxes, labs, scores)], losses


class LabelAwareMLPHead(nn.Module):
    def __init__(self, in_channels, representation_size):
        super().__init__()
        self.fc6 = nn.Linear(in_channels, representation_size)
        self.fc7 = nn.Linear(representation_size, representation_size)

    def forward(self, x, box_labels=None):
        return F.relu(self.fc7(F.relu(self.fc6(x.flatten(start_dim=1)))))


class RefFasterRCNN(GeneralizedRCNN):
    """fasterrcnn.py:354-499 + factory :307-329 with pretrained=False."""

    def __init__(self, num_classes=9, min_size=800, max_size=1333, box_batch_size_per_image=512,
                 rpn_batch_size_per_image=256, trainable_layers=5):
        backbone = resnet_fpn_backbone(backbone_name="resnet50", weights=None, trainable_layers=trainable_layers)
        oc = backbone.out_channels
        ag = AnchorGenerator(((32,), (64,), (128,), (256,), (512,)), ((0.5, 1.0, 2.0),) * 5)
        rpn = PerImageRPN(ag, RPNHead(oc, ag.num_anchors_per_location()[0]), 0.7, 0.3, rpn_batch_size_per_image, 0.5,
                          dict(training=2000, testing=1000), dict(training=2000, testing=1000), 0.7)
        roi = PerImageRoIHeads(MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2), LabelAwareMLPHead(oc * 49, 1024),
                               FastRCNNPredictor(1024, num_classes), 0.5, 0.5, box_batch_size_per_image, 0.25, None,
                               0.05, 0.5, 100)
        super().__init__(backbone, rpn, roi, GeneralizedRCNNTransform(min_size, max_size, [0.0] * 3, [1.0] * 3))

    def forward(self, images, targets=None):
        original = [tuple(img.shape[-2:]) for img in images]
        image_list, targets = self.transform(images, targets)
        features = self.backbone(image_list.tensors)
        proposals, rpn_losses = self.rpn(image_list, features, targets)
        detections, roi_losses = self.roi_heads(features, proposals, image_list.image_sizes, targets)
        detections = self.transform.postprocess(detections, image_list.image_sizes, original)
        for i, det in enumerate(detections):                             # fasterrcnn.py:492-497
            det["losses"] = {k: v[i] for k, v in {**rpn_losses, **roi_losses}.items()}
        return detections


class _Reverse(torch.autograd.Function):
    """DGcommon.py:33-45."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.neg() * 0.1


class _ImageHead(nn.Module):
    """DGcommon.py:48-81."""

    def __init__(self, n_dom):
        super().__init__()
        self.Conv1 = nn.Conv2d(256, 256, 3, stride=(2, 4))
        self.Conv2 = nn.Conv2d(256, 256, 3, stride=4)
        self.Conv3 = nn.Conv2d(256, 256, 3, stride=4)
        self.Conv4 = nn.Conv2d(256, 256, 3, stride=3)
        self.linear1 = nn.Linear(256, 128)
        self.linear2 = nn.Linear(128, n_dom)
        for c in (self.Conv1, self.Conv2, self.Conv3, self.Conv4):
            nn.init.normal_(c.weight, std=0.001)
            nn.init.constant_(c.bias, 0)

    def forward(self, x):
        x = _Reverse.apply(x)
        for c in (self.Conv1, self.Conv2, self.Conv3, self.Conv4):
            x = F.relu(c(x))
        return torch.sigmoid(self.linear2(F.relu(self.linear1(x.flatten(1)))))


class _InstanceHead(nn.Module):
    """DGFRCNN.py:4-64."""

    def __init__(self, n_out, reverse):
        super().__init__()
        self.dc_ip1, self.dc_ip2, self.classifer = nn.Linear(1024, 512), nn.Linear(512, 256), nn.Linear(256, n_out)
        self.reverse = reverse

    def forward(self, x):
        if self.reverse:
            x = _Reverse.apply(x)
        return torch.sigmoid(self.classifer(self.dc_ip2(F.relu(self.dc_ip1(x)))))


class RefDGFRCNN(nn.Module):
    """DGFRCNN.py:73-201 on the CPU (device-specific `.cuda()` lines dropped)."""

    CYCLE = (0, 1, 0, 2, 0, 3, 0, 4)

    def __init__(self, n_classes, batch_size, reg_weights, num_domains, min_size=600, max_size=1200):
        super().__init__()
        self.batch_size, self.reg_weights, self.num_domains = batch_size, list(reg_weights), num_domains
        self.InsDA = _InstanceHead(num_domains, True)
        self.InsClsPrime = nn.ModuleList([_InstanceHead(n_classes, True) for _ in range(num_domains)])
        self.InsCls = nn.ModuleList([_InstanceHead(n_classes, False) for _ in range(num_domains)])
        # DGFRCNN.py:81: fasterrcnn_resnet50_fpn(num_classes=n_classes, pretrained=True, trainable_backbone_layers=3)
        # -> FastRCNNPredictor(in_features, n_classes + 1) (fasterrcnn.py:327), conv1 / layer1 frozen (fasterrcnn.py:317)
        self.detector = RefFasterRCNN(n_classes + 1, min_size, max_size, trainable_layers=3)
        self.ImageDA = _ImageHead(num_domains)
        self.detector.backbone.register_forward_hook(lambda m, i, o: setattr(self, "base_feat", o))
        self.detector.roi_heads.box_head.register_forward_hook(self._grab)
        self.step_index = 0

    def _grab(self, module, inputs, output):
        self.box_features, self.box_labels = output, inputs[1]

    def configure_optimizer(self, lr=2e-3):
        return torch.optim.SGD([{"params": m.parameters(), "lr": lr, "weight_decay": 5e-4}
                                for m in (self.detector, self.ImageDA, self.InsDA, self.InsCls, self.InsClsPrime)])

    def training_step(self, batch):
        imgs, boxes, labels, domain = batch
        targets = [{"boxes": b.float(), "labels": l.long()} for b, l in zip(boxes, labels)]
        mode = self.CYCLE[self.step_index % len(self.CYCLE)]
        self.step_index += 1
        w = self.reg_weights
        if mode == 0:
            return sum(v for d in self.detector(imgs, targets) for v in d["losses"].values())
        if mode == 1:
            self.detector(imgs, targets)
            img_s = self.ImageDA(self.base_feat["0"])
            ida = self.InsDA(self.box_features)
            rep = int(ida.shape[0] / self.batch_size)
            ins_lab = domain.reshape(self.batch_size, 1).repeat(1, rep).reshape(ida.shape[0])
            return (w[0] * F.cross_entropy(img_s, domain) + w[1] * F.cross_entropy(ida, ins_lab)
                    + w[2] * F.mse_loss(ida, img_s.repeat(1, rep).reshape(ida.shape[0], self.num_domains)))
        for head in self.InsCls:
            for p in head.parameters():
                p.requires_grad = mode != 4
        per = []
        for i in range(len(imgs)):
            d = int(domain[i])
            if mode == 2:
                with torch.no_grad():
                    self.detector([imgs[i]], [targets[i]])
                per.append(F.cross_entropy(self.InsCls[d](self.box_features), self.box_labels[0]))
            elif mode == 3:
                self.detector([imgs[i]], [targets[i]])
                per.append(F.cross_entropy(self.InsClsPrime[d](self.box_features), self.box_labels[0]))
            else:
                self.detector([imgs[i]], [targets[i]])
                per.extend(F.cross_entropy(self.InsCls[j](self.box_features), self.box_labels[0])
                           for j in range(self.num_domains) if j != d)
        return (w[3] if mode == 3 else w[4]) * torch.mean(torch.stack(per))


def build_like_reference_factory(num_classes_with_bg: int, min_size: int, max_size: int) -> RefFasterRCNN:
    """Constructs the detector consuming the RNG exactly like fasterrcnn.fasterrcnn_resnet50_fpn
    (fasterrcnn.py:307-329): a 91-class FastWILDS first, then a fresh predictor — so that under the
    same torch.manual_seed the weights equal the reference's."""
    model = RefFasterRCNN(91, min_size, max_size)
    model.roi_heads.box_predictor = FastRCNNPredictor(1024, num_classes_with_bg)
    return model
