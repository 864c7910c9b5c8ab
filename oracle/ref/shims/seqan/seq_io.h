#pragma once
#include <seqan/sequence.h>
