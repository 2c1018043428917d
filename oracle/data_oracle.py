"""CPU restatement of the reference's per-sample input preparation (SURVEY f-2).  TEST INFRASTRUCTURE ONLY.

* ``decode_features``: ``get_img_feature`` of /root/reference/oscar/oscar_datasets_ml/oscar_tsv4.py:696-727
  (``np.frombuffer(base64.b64decode(text), dtype=np.float32).reshape(num_boxes, dim)`` -> ``torch.tensor(dtype)``)
  followed by the zero padding to ``max_img_seq_length`` of convert_example_to_features (:1050-1060).
* ``random_word_ids`` / ``random_phrases_ids``: ``random_word`` (:782-820) and ``random_phrases`` (:822-850) on token
  IDS, with the reference's ``random.random()`` / ``random.choice`` / ``random.randint`` draws replaced by supplied
  numbers (``u`` uniforms, ``r`` integer draws) so that the CUDA kernel can consume the identical values.
"""
import base64

import numpy as np
import torch


def decode_features(b64_texts, num_boxes, max_regions, feature_dim, dtype=torch.float32):
    out = torch.zeros(len(b64_texts), max_regions, feature_dim, dtype=dtype)
    for b, (text, nb) in enumerate(zip(b64_texts, num_boxes)):
        feat = np.frombuffer(base64.b64decode(text), dtype=np.float32).reshape((nb, feature_dim))  # :716-718
        feat = torch.tensor(np.copy(feat), dtype=dtype)                                            # :722
        n = min(nb, max_regions)
        out[b, :n] = feat[:n]
    return out


def random_word_ids(ids, u, r, mask_id, word_vocab):
    """ids: list of token ids (modified in place).  Returns labels (:782-820)."""
    labels = []
    for i, tok in enumerate(ids):
        prob = float(u[i])
        if prob < 0.15:
            prob /= 0.15
            if prob < 0.8:
                ids[i] = mask_id
            elif prob < 0.9:
                ids[i] = int(r[i]) % word_vocab
            labels.append(tok)
        else:
            labels.append(-1)
    return ids, labels


def random_phrases_ids(phrases, t1_label, phrase_mask_map, u, r, mask_id, phrase_vocab, vocab_size):
    """phrases: list of phrase ids (modified in place); its labels are discarded by the caller (:960)."""
    already = set()
    for i, t in enumerate(t1_label):
        if t >= 0 and i in phrase_mask_map:
            already.update(phrase_mask_map[i])
    for j in range(len(phrases)):
        if j in already:
            phrases[j] = mask_id
        else:
            prob = float(u[j])
            if prob < 0.15:
                prob /= 0.15
                if prob < 0.8:
                    phrases[j] = mask_id
                elif prob < 0.9:
                    phrases[j] = int(r[j]) % phrase_vocab + vocab_size
    return phrases
