"""CPU oracle for the DiffPool forward/backward hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch restatement of the reference's algorithm
(JiaxuanYou/graph-pooling, ``encoders.py:976-1334`` plus the commented-out DiffPool
``GraphConv`` at ``encoders.py:296-328`` == ``:945-974``, and ``set2set.py:8-57`` for the
base-set2set readout).  It exists so that the CUDA
path in ``graph_pooling_b200`` can be checked against it; it is NOT part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import it, and only as the checker / the CPU baseline.

Parity pin: ``tests/golden/make_golden.py`` imports the *actual* ``/root/reference/encoders.py``
(with the one class the shipped file lost -- R1 -- injected, ``.cuda()`` made a no-op -- R2 --
and the two un-runnable lines of ``loss()`` patched at run time -- R3/R4), runs it on seeded
inputs and stores inputs/weights/outputs/gradients in ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` checks this oracle against those vectors.  The reference ships
no tests or golden vectors of its own.

Repairs relative to the shipped text, each a deliberate decision:
  R1  DiffPool GraphConv restored from the commented-out code (encoders.py:296-328).
  R2  hard-coded .cuda() replaced by the input's device (encoders.py:1046,1051,1130,1317).
  R3  ``torch.min(pred_adj, torch.Tensor(1))`` (encoders.py:1317, an UNINITIALISED tensor)
      -> ``clamp(max=1.0)`` (the evident intent; inactive for adj_hop=1 because P<=1).
  R4  uint8 mask indexing (encoders.py:1329) -> multiply by the {0,1} mask.
  R5  per-level modules registered (the reference keeps levels < P-1 in plain lists only).
      Keys: the LAST level keeps the reference's attribute names (conv_first2, assign_conv_first,
      assign_pred ...) so a P=1 state-dict is identical to the reference's; level i < P-1 is
      registered as ``<name>_l{i}``.
  R6  forward uses the level's own assign_pred and the pooled feature width as the
      level>=1 assignment input dim (encoders.py:1273,1214).
  R7  link loss pairs level-0 S with the level-0 adjacency (encoders.py:1311 vs 1321).
  R8  behavioural quirks KEPT: SoftPooling ignores bn/dropout for the first GCN
      (encoders.py:1172-1173); BatchNorm is a fresh train-mode module on every call
      (encoders.py:1048-1052), i.e. batch statistics even in eval(); base path unmasked.
  R9  F.cross_entropy(size_average=True) -> reduction='mean' (encoders.py:1127).
  R10 init.xavier_uniform / init.constant -> the in-place variants (encoders.py:1005-1007).
  R11 sum of n_b^2 in int64 (encoders.py:1326).
North-star additions that the reference does NOT contain (parity unpinned by the reference,
oracle = the DiffPool paper's definitions): ``frobenius_link_loss`` and ``row_entropy_loss``.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

EPS_NORM = 1e-12   # F.normalize default              (encoders.py:326)
EPS_BN = 1e-5      # nn.BatchNorm1d default            (encoders.py:1051)
EPS_LINK = 1e-7    # eps in SoftPoolingGcnEncoder.loss (encoders.py:1307)


# ----------------------------------------------------------------------------------------
# functional pieces (used by the op-level parity tests)
# ----------------------------------------------------------------------------------------
def graph_conv(x, adj, weight, bias=None, add_self=False, normalize=True):
    """encoders.py:315-328: y = normalize((adj@x [+x]) @ W + b) over the feature dim."""
    y = torch.matmul(adj, x)
    if add_self:
        y = y + x
    y = torch.matmul(y, weight)
    if bias is not None:
        y = y + bias
    if normalize:
        y = F.normalize(y, p=2, dim=2)
    return y


def bn_per_node(x):
    """encoders.py:1048-1052: a fresh BatchNorm1d(num_nodes) in train mode: gamma=1, beta=0,
    biased batch variance, channel = node index, statistics over (batch, feature)."""
    mean = x.mean(dim=(0, 2), keepdim=True)
    var = x.var(dim=(0, 2), unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + EPS_BN)


def construct_mask(max_nodes, batch_num_nodes, device, dtype):
    """encoders.py:1035-1046: M[b, i, 0] = 1 if i < n_b."""
    n = torch.as_tensor(np.asarray(batch_num_nodes).astype(np.int64), device=device)
    idx = torch.arange(max_nodes, device=device)
    return (idx[None, :] < n[:, None]).to(dtype).unsqueeze(2)


def assign_softmax(sa, weight, bias, mask=None):
    """encoders.py:1273-1275: S = softmax(Linear(sa), -1) [* mask]."""
    s = torch.softmax(F.linear(sa, weight, bias), dim=-1)
    if mask is not None:
        s = s * mask
    return s


def pool(s, z, adj):
    """encoders.py:1278-1279: X' = S^T Z ; A' = (S^T A) S (left to right)."""
    st = torch.transpose(s, 1, 2)
    return torch.matmul(st, z), st @ adj @ s


def link_pred_loss(s, adj, batch_num_nodes=None, adj_hop=1):
    """encoders.py:1311-1331 (masked BCE between S S^T and A, normalised by sum n_b^2)."""
    pred_adj0 = s @ torch.transpose(s, 1, 2)
    tmp = pred_adj0
    pred_adj = pred_adj0
    for _ in range(adj_hop - 1):
        tmp = tmp @ pred_adj0
        pred_adj = pred_adj + tmp
    pred_adj = torch.clamp(pred_adj, max=1.0)                                   # R3
    ll = -adj * torch.log(pred_adj + EPS_LINK) - (1 - adj) * torch.log(1 - pred_adj + EPS_LINK)
    max_num_nodes = adj.size(1)
    if batch_num_nodes is None:
        num_entries = max_num_nodes * max_num_nodes * adj.size(0)
    else:
        nb = np.asarray(batch_num_nodes).astype(np.int64)                      # R11
        num_entries = int(np.sum(nb * nb))
        m = construct_mask(max_num_nodes, batch_num_nodes, adj.device, adj.dtype)
        ll = ll * (m @ torch.transpose(m, 1, 2))                               # R4
    return torch.sum(ll) / float(num_entries)


def frobenius_link_loss(s, adj, batch_num_nodes=None):
    """North-star option (NOT in the reference): mean over graphs of ||A - S S^T||_F taken over
    the real n_b x n_b block.  Parity unpinned by the reference."""
    p = s @ torch.transpose(s, 1, 2)
    d = adj - p
    if batch_num_nodes is not None:
        m = construct_mask(adj.size(1), batch_num_nodes, adj.device, adj.dtype)
        d = d * (m @ torch.transpose(m, 1, 2))
    return torch.sqrt((d * d).sum(dim=(1, 2))).mean()


def row_entropy_loss(s, batch_num_nodes=None, eps=1e-7):
    """North-star option (NOT in the reference): mean over real rows of -sum_k S log(S+eps)."""
    ent = -(s * torch.log(s + eps)).sum(dim=-1)
    if batch_num_nodes is None:
        return ent.mean()
    m = construct_mask(s.size(1), batch_num_nodes, s.device, s.dtype).squeeze(2)
    return (ent * m).sum() / float(int(np.sum(np.asarray(batch_num_nodes).astype(np.int64))))


# ----------------------------------------------------------------------------------------
# modules (same constructor signatures / parameter names as the reference)
# ----------------------------------------------------------------------------------------
class GraphConv(nn.Module):
    """R1: encoders.py:296-328."""

    def __init__(self, input_dim, output_dim, add_self=False, normalize_embedding=False,
                 dropout=0.0, bias=True):
        super().__init__()
        self.add_self = add_self
        self.dropout = dropout
        if dropout > 0.001:
            self.dropout_layer = nn.Dropout(p=dropout)
        self.normalize_embedding = normalize_embedding
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.weight = nn.Parameter(torch.empty(input_dim, output_dim))
        nn.init.xavier_uniform_(self.weight.data, gain=nn.init.calculate_gain('relu'))
        if bias:
            self.bias = nn.Parameter(torch.zeros(output_dim))
        else:
            self.bias = None

    def forward(self, x, adj):
        if self.dropout > 0.001:
            x = self.dropout_layer(x)
        return graph_conv(x, adj, self.weight, self.bias, self.add_self, self.normalize_embedding)


class GcnEncoderGraph(nn.Module):
    """encoders.py:976-1134."""

    def __init__(self, input_dim, hidden_dim, embedding_dim, label_dim, num_layers,
                 pred_hidden_dims=[], concat=True, bn=True, dropout=0.0, args=None):
        super().__init__()
        self.concat = concat
        add_self = not concat
        self.bn = bn
        self.num_layers = num_layers
        self.num_aggs = 1
        self.bias = True
        if args is not None:
            self.bias = args.bias
        self.conv_first, self.conv_block, self.conv_last = self.build_conv_layers(
            input_dim, hidden_dim, embedding_dim, num_layers, add_self, normalize=True, dropout=dropout)
        self.act = nn.ReLU()
        self.label_dim = label_dim
        if concat:
            self.pred_input_dim = hidden_dim * (num_layers - 1) + embedding_dim
        else:
            self.pred_input_dim = embedding_dim
        self.pred_model = self.build_pred_layers(self.pred_input_dim, pred_hidden_dims, label_dim,
                                                 num_aggs=self.num_aggs)
        self._reinit()

    def _reinit(self):
        for m in self.modules():                                              # encoders.py:1003-1007
            if isinstance(m, GraphConv):
                nn.init.xavier_uniform_(m.weight.data, gain=nn.init.calculate_gain('relu'))
                if m.bias is not None:
                    nn.init.constant_(m.bias.data, 0.0)

    def build_conv_layers(self, input_dim, hidden_dim, embedding_dim, num_layers, add_self,
                          normalize=False, dropout=0.0):
        conv_first = GraphConv(input_dim=input_dim, output_dim=hidden_dim, add_self=add_self,
                               normalize_embedding=normalize, bias=self.bias)
        conv_block = nn.ModuleList(
            [GraphConv(input_dim=hidden_dim, output_dim=hidden_dim, add_self=add_self,
                       normalize_embedding=normalize, dropout=dropout, bias=self.bias)
             for _ in range(num_layers - 2)])
        conv_last = GraphConv(input_dim=hidden_dim, output_dim=embedding_dim, add_self=add_self,
                              normalize_embedding=normalize, bias=self.bias)
        return conv_first, conv_block, conv_last

    def build_pred_layers(self, pred_input_dim, pred_hidden_dims, label_dim, num_aggs=1):
        pred_input_dim = pred_input_dim * num_aggs
        if len(pred_hidden_dims) == 0:
            return nn.Linear(pred_input_dim, label_dim)
        layers = []
        for pred_dim in pred_hidden_dims:
            layers.append(nn.Linear(pred_input_dim, pred_dim))
            layers.append(self.act)
            pred_input_dim = pred_dim
        layers.append(nn.Linear(pred_dim, label_dim))
        return nn.Sequential(*layers)

    def construct_mask(self, max_nodes, batch_num_nodes, like):
        return construct_mask(max_nodes, batch_num_nodes, like.device, like.dtype)

    def apply_bn(self, x):
        return bn_per_node(x)

    def gcn_forward(self, x, adj, conv_first, conv_block, conv_last, embedding_mask=None):
        x = conv_first(x, adj)
        x = self.act(x)
        if self.bn:
            x = self.apply_bn(x)
        x_all = [x]
        for i in range(len(conv_block)):
            x = conv_block[i](x, adj)
            x = self.act(x)
            if self.bn:
                x = self.apply_bn(x)
            x_all.append(x)
        x = conv_last(x, adj)
        x_all.append(x)
        x_tensor = torch.cat(x_all, dim=2)
        if embedding_mask is not None:
            x_tensor = x_tensor * embedding_mask
        return x_tensor

    def forward(self, x, adj, batch_num_nodes=None, **kwargs):
        # encoders.py:1083-1122 -- the mask is built (:1087) but never applied on this path.
        x = self.conv_first(x, adj)
        x = self.act(x)
        if self.bn:
            x = self.apply_bn(x)
        out_all = [torch.max(x, dim=1)[0]]
        for i in range(self.num_layers - 2):
            x = self.conv_block[i](x, adj)
            x = self.act(x)
            if self.bn:
                x = self.apply_bn(x)
            out_all.append(torch.max(x, dim=1)[0])
        x = self.conv_last(x, adj)
        out = torch.max(x, dim=1)[0]
        out_all.append(out)
        output = torch.cat(out_all, dim=1) if self.concat else out
        return self.pred_model(output)

    def loss(self, pred, label, type='softmax'):
        if type == 'softmax':
            return F.cross_entropy(pred, label, reduction='mean')             # R9
        elif type == 'margin':
            onehot = torch.zeros(pred.size(0), self.label_dim, dtype=torch.long, device=pred.device)
            onehot.scatter_(1, label.view(-1, 1), 1)
            return torch.nn.MultiLabelMarginLoss()(pred, onehot)


class Set2Set(nn.Module):
    """set2set.py:8-57 (R2: the zero states follow the input's device / dtype instead of .cuda())."""

    def __init__(self, input_dim, hidden_dim, act_fn=nn.ReLU, num_layers=1):
        super().__init__()
        self.input_dim = input_dim
        self.hidden_dim = hidden_dim
        self.num_layers = num_layers
        if hidden_dim <= input_dim:
            print('ERROR: Set2Set output_dim should be larger than input_dim')
        self.lstm_output_dim = hidden_dim - input_dim
        self.lstm = nn.LSTM(hidden_dim, input_dim, num_layers=num_layers, batch_first=True)
        self.pred = nn.Linear(hidden_dim, input_dim)
        self.act = act_fn()

    def forward(self, embedding):
        batch_size, n = embedding.size(0), embedding.size(1)
        z = lambda *shape: torch.zeros(*shape, dtype=embedding.dtype, device=embedding.device)
        hidden = (z(self.num_layers, batch_size, self.lstm_output_dim),
                  z(self.num_layers, batch_size, self.lstm_output_dim))
        q_star = z(batch_size, 1, self.hidden_dim)
        for _ in range(n):                                                    # set2set.py:47-55
            q, hidden = self.lstm(q_star, hidden)
            e = embedding @ torch.transpose(q, 1, 2)
            a = torch.softmax(e, dim=1)
            r = torch.sum(a * embedding, dim=1, keepdim=True)
            q_star = torch.cat((q, r), dim=2)
        q_star = torch.squeeze(q_star, dim=1)
        return self.act(self.pred(q_star))


class GcnSet2SetEncoder(GcnEncoderGraph):
    """encoders.py:1137-1157 (method=base-set2set): masked GCN concat -> Set2Set readout -> pred_model."""

    def __init__(self, input_dim, hidden_dim, embedding_dim, label_dim, num_layers,
                 pred_hidden_dims=[], concat=True, bn=True, dropout=0.0, args=None):
        super().__init__(input_dim, hidden_dim, embedding_dim, label_dim, num_layers, pred_hidden_dims, concat,
                         bn, dropout, args=args)
        self.s2s = Set2Set(self.pred_input_dim, self.pred_input_dim * 2)

    def forward(self, x, adj, batch_num_nodes=None, **kwargs):
        mask = None
        if batch_num_nodes is not None:
            mask = self.construct_mask(adj.size(1), batch_num_nodes, x)
        emb = self.gcn_forward(x, adj, self.conv_first, self.conv_block, self.conv_last, mask)
        return self.pred_model(self.s2s(emb))


class SoftPoolingGcnEncoder(GcnEncoderGraph):
    """encoders.py:1160-1334."""

    def __init__(self, max_num_nodes, input_dim, hidden_dim, embedding_dim, label_dim, num_layers,
                 assign_hidden_dim, assign_ratio=0.25, assign_num_layers=-1, num_pooling=1,
                 pred_hidden_dims=[50], concat=True, bn=True, dropout=0.0, linkpred=True,
                 assign_input_dim=-1, args=None):
        # R8: bn / dropout are NOT forwarded to the first GCN (encoders.py:1172-1173).
        super().__init__(input_dim, hidden_dim, embedding_dim, label_dim, num_layers,
                         pred_hidden_dims=pred_hidden_dims, concat=concat, args=args)
        add_self = not concat
        self.num_pooling = num_pooling
        self.linkpred = linkpred
        self.assign_ent = True

        def reg(name, i, mod):                                                # R5
            setattr(self, name if i == num_pooling - 1 else '%s_l%d' % (name, i), mod)
            return mod

        self.conv_first_after_pool, self.conv_block_after_pool, self.conv_last_after_pool = [], [], []
        for i in range(num_pooling):
            f, b, l = self.build_conv_layers(self.pred_input_dim, hidden_dim, embedding_dim, num_layers,
                                             add_self, normalize=True, dropout=dropout)
            self.conv_first_after_pool.append(reg('conv_first2', i, f))
            self.conv_block_after_pool.append(reg('conv_block2', i, b))
            self.conv_last_after_pool.append(reg('conv_last2', i, l))

        if assign_num_layers == -1:
            assign_num_layers = num_layers
        if assign_input_dim == -1:
            assign_input_dim = input_dim
        self.assign_conv_first_modules, self.assign_conv_block_modules = [], []
        self.assign_conv_last_modules, self.assign_pred_modules = [], []
        self.assign_dims = []
        assign_dim = int(max_num_nodes * assign_ratio)
        for i in range(num_pooling):
            self.assign_dims.append(assign_dim)
            f, b, l = self.build_conv_layers(assign_input_dim, assign_hidden_dim, assign_dim,
                                             assign_num_layers, add_self, normalize=True)
            apin = assign_hidden_dim * (num_layers - 1) + assign_dim if concat else assign_dim
            ap = self.build_pred_layers(apin, [], assign_dim, num_aggs=1)
            assign_input_dim = self.pred_input_dim                            # R6 (reference: embedding_dim)
            assign_dim = int(assign_dim * assign_ratio)
            self.assign_conv_first_modules.append(reg('assign_conv_first', i, f))
            self.assign_conv_block_modules.append(reg('assign_conv_block', i, b))
            self.assign_conv_last_modules.append(reg('assign_conv_last', i, l))
            self.assign_pred_modules.append(reg('assign_pred', i, ap))

        self.pred_model = self.build_pred_layers(self.pred_input_dim * (num_pooling + 1), pred_hidden_dims,
                                                 label_dim, num_aggs=self.num_aggs)
        self._reinit()

    def forward(self, x, adj, batch_num_nodes, **kwargs):
        x_a = kwargs['assign_x'] if 'assign_x' in kwargs else x
        max_num_nodes = adj.size(1)
        mask0 = None
        if batch_num_nodes is not None:
            mask0 = self.construct_mask(max_num_nodes, batch_num_nodes, x)
        out_all = []
        embedding_tensor = self.gcn_forward(x, adj, self.conv_first, self.conv_block, self.conv_last, mask0)
        out_all.append(torch.max(embedding_tensor, dim=1)[0])
        self.assign_tensors = []
        self.pooled = []
        for i in range(self.num_pooling):
            embedding_mask = mask0 if i == 0 else None
            a = self.gcn_forward(x_a, adj, self.assign_conv_first_modules[i],
                                 self.assign_conv_block_modules[i], self.assign_conv_last_modules[i],
                                 embedding_mask)
            ap = self.assign_pred_modules[i]                                  # R6
            self.assign_tensor = assign_softmax(a, ap.weight, ap.bias, embedding_mask)
            self.assign_tensors.append(self.assign_tensor)
            x, adj = pool(self.assign_tensor, embedding_tensor, adj)
            self.pooled.append((x, adj))
            x_a = x
            embedding_tensor = self.gcn_forward(x, adj, self.conv_first_after_pool[i],
                                                self.conv_block_after_pool[i], self.conv_last_after_pool[i])
            out = torch.max(embedding_tensor, dim=1)[0]
            out_all.append(out)
        output = torch.cat(out_all, dim=1) if self.concat else out
        return self.pred_model(output)

    def loss(self, pred, label, adj=None, batch_num_nodes=None, adj_hop=1):
        loss = super().loss(pred, label)
        if self.linkpred:
            s0 = self.assign_tensors[0]                                       # R7
            self.link_loss = link_pred_loss(s0, adj, batch_num_nodes, adj_hop)
            return loss + self.link_loss
        return loss


def train_step(model, x, adj, label, batch_num_nodes, assign_x=None, linkpred=True, optimizer=None,
               clip=2.0):
    """train.py:196-210: zero_grad -> forward -> loss -> backward [-> clip -> Adam]."""
    model.zero_grad()
    kw = {} if assign_x is None else {'assign_x': assign_x}
    ypred = model(x, adj, batch_num_nodes, **kw)
    if isinstance(model, SoftPoolingGcnEncoder) and linkpred and model.linkpred:
        loss = model.loss(ypred, label, adj, batch_num_nodes)
    else:
        loss = model.loss(ypred, label)
    loss.backward()
    if optimizer is not None:
        nn.utils.clip_grad_norm_(model.parameters(), clip)
        optimizer.step()
    return ypred, loss
