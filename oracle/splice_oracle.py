"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's embedding splice and conditioning gather.

`splice` follows DreamLLMModel.forward, modeling_dreamllm.py:1082-1099 (dream queries) and :1104-1141 (image features):
per-sample torch.where + torch.cat, images counted globally across the batch (`cur_image_idx`), extra <im_start> tokens
beyond the number of images left untouched (:1122-1123).  `gather_conditioning` follows DreamLLMForCausalMLM.forward
:1401-1418.  Pinned against the live reference in tests/test_oracle_pin.py when /root/reference is present."""
import torch


def splice(input_ids, inputs_embeds, image_features, dream_queries, image_start_id, dream_start_id):
    B = input_ids.shape[0]
    if dream_queries is not None:
        query_embedding = dream_queries.repeat(B, 1, 1)                               # DreamEmbedding.forward, plugins:180-181
        Q = dream_queries.shape[1]
        new = []
        for cur_ids, cur, cur_q in zip(input_ids, inputs_embeds, query_embedding):
            pos = torch.where(cur_ids == dream_start_id)[0]
            out = cur
            for p in pos:
                out = torch.cat([out[: p + 1], cur_q.to(out.dtype), out[p + Q + 1:]], dim=0)
            assert out.shape[0] == cur.shape[0]
            new.append(out)
        inputs_embeds = torch.stack(new, dim=0)
    if image_features is not None:
        new = []
        idx = 0
        for cur_ids, cur in zip(input_ids, inputs_embeds):
            pos = torch.where(cur_ids == image_start_id)[0]
            out = cur
            for p in pos:
                if idx >= image_features.shape[0]:
                    break
                f = image_features[idx]
                n = f.shape[0]
                assert p + n + 1 <= cur.shape[0]
                out = torch.cat((out[: p + 1], f.to(out.dtype), out[p + n + 1:]), dim=0)
                idx += 1
            assert out.shape[0] == cur.shape[0]
            new.append(out)
        inputs_embeds = torch.stack(new, dim=0)
    return inputs_embeds


def gather_conditioning(input_ids, last_hidden, dream_start_id, Q, n_dm):
    rows = []
    for cur_ids, cur_h in zip(input_ids, last_hidden):
        for p in torch.where(cur_ids == dream_start_id)[0]:
            if len(rows) >= n_dm:
                break
            rows.append(cur_h[p + 1: p + 1 + Q])
    return torch.stack(rows, dim=0)
