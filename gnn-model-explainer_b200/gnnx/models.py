"""Drop-in parameter containers for the reference's models.py (GraphConv, GcnEncoderGraph,
GcnEncoderNode): same constructor signatures, same parameter names / state_dict keys
(conv_first.weight, conv_block.i.weight, conv_last.weight, pred_model.weight, ...), same
initialisation (models.py:136-150), so reference checkpoints load with load_state_dict.

The explainer never calls these forwards: libgnnx's persistent kernel carries its own fused
forward/backward of exactly this architecture (models.py:58-80,230-267,363-376).  The torch
`forward` below exists for API completeness (prediction outside the explainer) and as the
plain-PyTorch fp32 statement of the op that the kernel's forward is tested against."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import init


class GraphConv(nn.Module):
    """models.py:9-80.  y = normalize((adj @ x) @ W + b); att / add_self variants are out of scope."""

    def __init__(self, input_dim, output_dim, add_self=False, normalize_embedding=False, dropout=0.0,
                 bias=True, gpu=True, att=False):
        super().__init__()
        if att or add_self:
            raise NotImplementedError("att / add_self GraphConv variants are out of scope (SURVEY 8f)")
        self.att = att
        self.add_self = add_self
        self.dropout = dropout
        if dropout > 0.001:
            self.dropout_layer = nn.Dropout(p=dropout)
        self.normalize_embedding = normalize_embedding
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.weight = nn.Parameter(torch.empty(input_dim, output_dim))
        self.bias = nn.Parameter(torch.empty(output_dim)) if bias else None

    def forward(self, x, adj):
        if self.dropout > 0.001:
            x = self.dropout_layer(x)
        y = torch.matmul(torch.matmul(adj, x), self.weight)
        if self.bias is not None:
            y = y + self.bias
        if self.normalize_embedding:
            y = F.normalize(y, p=2, dim=2)
        return y, adj


class GcnEncoderGraph(nn.Module):
    """models.py:83-329 (graph classification: per-layer max-pool readout)."""

    def __init__(self, input_dim, hidden_dim, embedding_dim, label_dim, num_layers, pred_hidden_dims=[],
                 concat=True, bn=True, dropout=0.0, add_self=False, args=None):
        super().__init__()
        if len(pred_hidden_dims) != 0:
            raise NotImplementedError("pred_hidden_dims != [] is not built")
        self.concat = concat
        self.bn = bn
        self.num_layers = num_layers
        self.num_aggs = 1
        self.bias = True if args is None else getattr(args, "bias", True)
        self.gpu = False if args is None else getattr(args, "gpu", False)
        self.att = (args is not None and getattr(args, "method", "base") == "att")
        if self.att:
            raise NotImplementedError("method='att' is out of scope (SURVEY 8f)")
        self.conv_first = GraphConv(input_dim, hidden_dim, add_self, True, 0.0, self.bias)
        self.conv_block = nn.ModuleList(
            [GraphConv(hidden_dim, hidden_dim, add_self, True, dropout, self.bias) for _ in range(num_layers - 2)])
        self.conv_last = GraphConv(hidden_dim, embedding_dim, add_self, True, 0.0, self.bias)
        self.act = nn.ReLU()
        self.label_dim = label_dim
        self.pred_input_dim = hidden_dim * (num_layers - 1) + embedding_dim if concat else embedding_dim
        self.pred_model = nn.Linear(self.pred_input_dim, label_dim)
        for m in self.modules():
            if isinstance(m, GraphConv):
                init.xavier_uniform_(m.weight.data, gain=nn.init.calculate_gain("relu"))
                if m.bias is not None:
                    init.constant_(m.bias.data, 0.0)

    def apply_bn(self, x):
        bn_module = nn.BatchNorm1d(x.size()[1]).to(x.device)
        return bn_module(x)

    def _layers(self, x, adj):
        outs = []
        x, _ = self.conv_first(x, adj)
        x = self.act(x)
        if self.bn:
            x = self.apply_bn(x)
        outs.append(x)
        for conv in self.conv_block:
            x, _ = conv(x, adj)
            x = self.act(x)
            if self.bn:
                x = self.apply_bn(x)
            outs.append(x)
        x, _ = self.conv_last(x, adj)
        outs.append(x)
        return outs

    def forward(self, x, adj, batch_num_nodes=None, **kwargs):
        outs = self._layers(x, adj)
        pooled = [torch.max(o, dim=1)[0] for o in outs]
        output = torch.cat(pooled, dim=1) if self.concat else pooled[-1]
        self.embedding_tensor = output
        adj_att = torch.stack([adj] * len(outs), dim=3)
        return self.pred_model(output), adj_att

    def loss(self, pred, label, type="softmax"):
        return F.cross_entropy(pred, label)


class GcnEncoderNode(GcnEncoderGraph):
    """models.py:331-380 (node classification: Linear over the concatenated per-layer embeddings)."""

    def __init__(self, input_dim, hidden_dim, embedding_dim, label_dim, num_layers, pred_hidden_dims=[],
                 concat=True, bn=True, dropout=0.0, args=None):
        super().__init__(input_dim, hidden_dim, embedding_dim, label_dim, num_layers, pred_hidden_dims,
                         concat, bn, dropout, args=args)
        self.celoss = nn.CrossEntropyLoss()

    def forward(self, x, adj, batch_num_nodes=None, **kwargs):
        outs = self._layers(x, adj)
        self.embedding_tensor = torch.cat(outs, dim=2) if self.concat else outs[-1]
        adj_att = torch.stack([adj] * len(outs), dim=3)
        return self.pred_model(self.embedding_tensor), adj_att

    def loss(self, pred, label):
        return self.celoss(torch.transpose(pred, 1, 2), label)
