from .sharded import ShardedSlidingWindowInferer, ShardPlan, exchange_partials, make_shard_plan
