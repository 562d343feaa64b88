from .attention import CausalLinearAttention


class AttentionBuilder:
    def __init__(self, **kw):
        self.kw = kw

    @classmethod
    def from_kwargs(cls, **kw):
        return cls(**kw)

    def get(self, attention_type):
        if attention_type != "causal-linear":
            raise ValueError("stand-in only provides causal-linear")
        return CausalLinearAttention(self.kw["query_dimensions"], feature_map=self.kw["feature_map"])
