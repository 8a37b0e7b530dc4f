"""Model families: Qwen3 dense / MoE (reference parity) plus Llama-3 and Mixtral assembled from the same blocks."""
