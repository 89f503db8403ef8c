"""Per-model hook trios (``llm_forward / decoder_forward / attention_forward``) — the reference keeps them under
``framefusion/models/<family>/``; see ``framefusion_b200.interface`` for how they are installed."""
