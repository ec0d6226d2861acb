"""Literal restatement of the reference TF graph, differentiated by torch
autograd -- the second, independent restatement the closed-form oracle is
checked against.  TEST INFRASTRUCTURE ONLY; tiny sizes only (the one-hot
matmuls of ``Write_Memory`` materialise [B, U, 4D] like the reference does).

Every statement mirrors one line of
``/root/reference/Code/Recommender/Model_Recommender.py`` (cited inline); the
gradient slices come from autograd on the gathered tensors (``retain_grad`` on
``P[u]`` / ``R[i]`` = the un-deduplicated IndexedSlices values TF produces).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def literal_step(P, R, Cat, G, feed, hyper, dtype=torch.float64):
    t = lambda x: torch.tensor(np.asarray(x), dtype=dtype)
    P, R, Cat, G = t(P).requires_grad_(), t(R).requires_grad_(), t(Cat).requires_grad_(), t(G)
    U, _, D = P.shape
    L = G.shape[0]
    user_input = torch.tensor(np.asarray(feed["user_input"]).astype(np.int64))
    item_input = torch.tensor(np.asarray(feed["item_input"]).astype(np.int64))
    labels = t(feed["labels"]).reshape(-1)
    write_sign = t(feed["write_sign"]).reshape(-1, 1)
    categories = t(feed["categories"]).reshape(-1, 4, 1)
    user_one_hot_label = t(feed["user_one_hot_label"]).reshape(-1, L)
    a = torch.tensor(hyper.high_level_score_coefficient, dtype=dtype)

    # ---- inference (:56-97)
    User_Memory = P[user_input]; User_Memory.retain_grad()                    # :57
    high_mem, low_mem = torch.split(User_Memory, [1, 4], dim=1)               # :59
    Item_Embedding_g = R[item_input]; Item_Embedding_g.retain_grad()          # :63
    Item_Embedding = Item_Embedding_g.unsqueeze(1)                            # :65
    Dish_Category = categories * Cat                                          # :67
    category_score = high_mem * Dish_Category                                 # :71
    rs_cat = category_score.sum(dim=(1, 2))                                   # :75
    category_num = categories.sum(dim=(1, 2))                                 # :77
    high_score = rs_cat / category_num                                        # :79
    Dish_Memory = categories * low_mem                                        # :82
    dish_score = Item_Embedding * Dish_Memory                                 # :86
    low_score = dish_score.sum(dim=(1, 2)) / category_num                     # :90-92
    score = a * high_score + (1 - a) * low_score                              # :95-96

    # ---- loss (:99-104)
    loss = F.binary_cross_entropy_with_logits(score, labels, reduction="mean")
    loss.backward()
    out = dict(scores=score.detach().numpy(), loss=float(loss.detach()),
               dP_slices=User_Memory.grad.numpy(), dR_slices=Item_Embedding_g.grad.numpy(),
               dCat=Cat.grad.numpy())
    sq = (out["dP_slices"] ** 2).sum() + (out["dR_slices"] ** 2).sum() + (out["dCat"] ** 2).sum()
    out["global_norm"] = float(np.sqrt(sq))                                   # clip_ops.global_norm

    # ---- Write_Memory (:106-220), reading the pre-step tables
    with torch.no_grad():
        item_embedding = R[item_input].unsqueeze(1)                           # :107-109
        dish_memory = categories * item_embedding                             # :111
        low_coefficient = (hyper.beta_1 * write_sign).unsqueeze(1)            # :115-117
        dish_memory = dish_memory * low_coefficient                           # :119
        dish_category = (categories * Cat).sum(dim=1)                         # :124-128
        cnum = categories.sum(dim=(1, 2)).unsqueeze(1)                        # :130-132
        dish_category = (dish_category / cnum).unsqueeze(1)                   # :134-138
        high_coefficient = (hyper.beta_2 * write_sign).unsqueeze(1)           # :140-142
        dish_category = dish_category * high_coefficient                      # :144
        dish_memory = dish_memory.reshape(-1, 4 * D).unsqueeze(1)             # :149
        user_onehot = F.one_hot(user_input, U).to(dtype).unsqueeze(2)         # :151
        dish_bias = torch.matmul(user_onehot, dish_memory).sum(0).reshape(-1, 4, D)      # :154-156
        category_bias = torch.matmul(user_onehot, dish_category).sum(0).unsqueeze(1)     # :158-162
        bias = torch.cat([category_bias, dish_bias], 1)                       # :165
        P1 = P + bias                                                         # :167
        label_onehot = user_one_hot_label.unsqueeze(2)                        # :170
        gm = G.reshape(-1, 5 * D)                                             # :174
        ulm = (label_onehot * gm).sum(dim=1)                                  # :176-180
        ulm = (ulm / user_one_hot_label.sum(dim=1, keepdim=True)).unsqueeze(1)  # :182-188
        general_bias = torch.matmul(user_onehot, ulm).sum(0).reshape(-1, 5, D)  # :190-194
        P2 = P1 + hyper.alpha * general_bias                                  # :196-198
        dgb = torch.matmul(label_onehot, dish_memory).sum(0).reshape(-1, 4, D)  # :201-205
        cgb = torch.matmul(label_onehot, dish_category).sum(0).unsqueeze(1)   # :207-211
        G1 = G + torch.cat([cgb, dgb], 1)                                     # :213-215
    out["P_after_write"] = P2.numpy()
    out["G_after_write"] = G1.numpy()
    out["personal"] = float(P2.mean())                                        # :218
    out["general"] = float(G1.mean())                                         # :219
    return out
