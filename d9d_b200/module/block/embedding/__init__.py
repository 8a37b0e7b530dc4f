from .shard_token_embedding import SplitTokenEmbeddings

__all__ = ["SplitTokenEmbeddings"]
